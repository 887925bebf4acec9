#!/usr/bin/env python
"""Times uc2_gemm_bf16 on the encoder's GEMM shapes (M = tokens per GPU of the bench workload) with the
fused epilogue each call site uses.  CUDA events around `reps` back-to-back launches per shape.

    python scripts/gemm_bench.py [--tokens 19200] [--reps 20]
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from uc2_b200 import _lib  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tokens", type=int, default=19200)
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--block-n", type=int, default=0)
    ap.add_argument("--ctas", type=int, default=0)
    ap.add_argument("--warm", type=int, default=3)
    a = ap.parse_args()
    M, H, F, Q = a.tokens, 768, 3072, 2304
    dev = "cuda"
    bf, f32 = torch.bfloat16, torch.float32
    r = lambda *s: (torch.randn(*s, device=dev) * 0.05).to(bf)
    x, w_qkv, w_o, w1, w2 = r(M, H), r(Q, H), r(H, H), r(F, H), r(H, F)
    g3072, dqkv = r(M, F), r(M, Q)
    bias = {n: torch.zeros(n, device=dev) for n in (H, F, Q)}
    res32 = torch.randn(M, H, device=dev)
    o768, o2304, o3072, pre3072 = (torch.empty(M, n, dtype=bf, device=dev) for n in (H, Q, F, F))
    z32 = torch.empty(M, H, dtype=f32, device=dev)
    dw = {k: torch.zeros(s, dtype=f32, device=dev) for k, s in
          dict(qkv=(Q, H), o=(H, H), f1=(F, H), f2=(H, F)).items()}
    bn = a.block_n
    import functools
    _lib.gemm = functools.partial(_lib.gemm, ctas=a.ctas)
    cases = [
        ("fwd QKV      bias                 ", M, Q, H, lambda: _lib.gemm(x, w_qkv, M, Q, H, bias=bias[Q], out_bf16=o2304, block_n=bn)),
        ("fwd O-proj   bias+res32 -> f32    ", M, H, H, lambda: _lib.gemm(x, w_o, M, H, H, bias=bias[H], residual=res32, out_f32=z32, block_n=bn)),
        ("fwd FFN1     bias+gelu (+pre)     ", M, F, H, lambda: _lib.gemm(x, w1, M, F, H, bias=bias[F], act=_lib.ACT_GELU, out_bf16=o3072, out_pre=pre3072, block_n=bn)),
        ("fwd FFN2     bias+res32 -> f32    ", M, H, F, lambda: _lib.gemm(g3072, w2, M, H, F, bias=bias[H], residual=res32, out_f32=z32, block_n=bn)),
        ("dgrad FFN2   dgelu(aux)           ", M, F, H, lambda: _lib.gemm(x, w2, M, F, H, b_mn=True, aux=pre3072, act=_lib.ACT_DGELU, out_bf16=o3072, block_n=bn)),
        ("dgrad FFN1   +res bf16            ", M, H, F, lambda: _lib.gemm(g3072, w1, M, H, F, b_mn=True, residual=x, out_bf16=o768, block_n=bn)),
        ("dgrad O-proj                      ", M, H, H, lambda: _lib.gemm(x, w_o, M, H, H, b_mn=True, out_bf16=o768, block_n=bn)),
        ("dgrad QKV    +res bf16            ", M, H, Q, lambda: _lib.gemm(dqkv, w_qkv, M, H, Q, b_mn=True, residual=x, out_bf16=o768, block_n=bn)),
        ("wgrad FFN2   [768,3072] K=tokens  ", H, F, M, lambda: _lib.gemm(x, g3072, H, F, M, a_mn=True, b_mn=True, out_f32=dw["f2"], accumulate=True, split_k=0, block_n=bn)),
        ("wgrad FFN1   [3072,768]           ", F, H, M, lambda: _lib.gemm(g3072, x, F, H, M, a_mn=True, b_mn=True, out_f32=dw["f1"], accumulate=True, split_k=0, block_n=bn)),
        ("wgrad O-proj [768,768]            ", H, H, M, lambda: _lib.gemm(x, x, H, H, M, a_mn=True, b_mn=True, out_f32=dw["o"], accumulate=True, split_k=0, block_n=bn)),
        ("wgrad QKV    [2304,768]           ", Q, H, M, lambda: _lib.gemm(dqkv, x, Q, H, M, a_mn=True, b_mn=True, out_f32=dw["qkv"], accumulate=True, split_k=0, block_n=bn)),
    ]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    tot_f = tot_ms = 0.0
    for name, m, n, k, fn in cases:
        for _ in range(a.warm):
            fn()
        torch.cuda.synchronize()
        # back-to-back launches (as inside a training step: launch latency hidden, the 150-300 MB working set
        # of one call already exceeds the 126 MB L2)
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / a.reps
        fl = 2.0 * m * n * k
        tot_f += fl; tot_ms += ms
        print(f"{name} M={m:6d} N={n:5d} K={k:6d}  {ms * 1e3:8.1f} us  {fl / ms / 1e9:7.1f} TFLOP/s")
    print(f"one layer fwd+bwd GEMMs: {tot_ms * 1e3:.1f} us, {tot_f / tot_ms / 1e9:.1f} TFLOP/s")


if __name__ == "__main__":
    main()
