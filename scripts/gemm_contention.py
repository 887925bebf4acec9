#!/usr/bin/env python
"""What the persistent GEMM loses when another kernel holds SMs (a stand-in for a NCCL collective: uc2_debug_occupy_sms),
with the fixed round-robin tile order and with tiles drawn from a device counter (uc2_gemm_sched_dynamic).

    python scripts/gemm_contention.py [--tokens 10240]
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from uc2_b200 import _lib  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tokens", type=int, default=10240)
    ap.add_argument("--reps", type=int, default=12)
    a = ap.parse_args()
    L = _lib.lib()
    M, H, F, Q = a.tokens, 768, 3072, 2304
    bf = torch.bfloat16
    r = lambda *s: (torch.randn(*s, device="cuda") * 0.05).to(bf)
    x, w_qkv, w1, w2, g = r(M, H), r(Q, H), r(F, H), r(H, F), r(M, F)
    bias = {n: torch.zeros(n, device="cuda") for n in (H, F, Q)}
    o2304, o3072, pre = (torch.empty(M, n, dtype=bf, device="cuda") for n in (Q, F, F))
    o768 = torch.empty(M, H, dtype=bf, device="cuda")
    cases = [("fwd QKV  K= 768", lambda: _lib.gemm(x, w_qkv, M, Q, H, bias=bias[Q], out_bf16=o2304), 2.0 * M * Q * H),
             ("fwd FFN1 K= 768", lambda: _lib.gemm(x, w1, M, F, H, bias=bias[F], act=_lib.ACT_GELU, out_bf16=o3072, out_pre=pre), 2.0 * M * F * H),
             ("dgrad FFN1 K=3072", lambda: _lib.gemm(g, w1, M, H, F, b_mn=True, residual=x, out_bf16=o768), 2.0 * M * F * H)]
    side = torch.cuda.Stream()
    print(f"{'GEMM':18s} {'SMs held':>8s} {'fixed order':>14s} {'counter':>14s}   (us per launch, {a.reps} back-to-back launches)")
    for name, fn, fl in cases:
        for held in (0, 16, 32, 64):
            row = []
            for dyn in (0, 1):
                L.uc2_gemm_sched_dynamic(dyn)
                for _ in range(3):
                    fn()
                torch.cuda.synchronize()
                if held:
                    L.uc2_debug_occupy_sms(held, 20_000_000, side.cuda_stream)      # ~10 ms
                    torch.cuda._sleep(200_000)                                       # let it become resident
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(a.reps):
                    fn()
                e1.record()
                torch.cuda.synchronize()
                row.append(e0.elapsed_time(e1) / a.reps * 1e3)
            print(f"{name:18s} {held:8d} {row[0]:14.1f} {row[1]:14.1f}")
    L.uc2_gemm_sched_dynamic(0)


if __name__ == "__main__":
    main()
