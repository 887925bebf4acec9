#!/usr/bin/env python
"""Attention fwd/bwd timing + check against a torch fp32 reference (key-padding mask, fixed list of shapes).
With UC2_ATTN_TCGEN05=1 in the environment the public entry points route S <= 160 to the experimental tcgen05
kernels (csrc/attention_tc.cu), so the same script times those."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from uc2_b200._lib import call, stream

def t(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3

for B, S in [(120, 160), (8, 76), (48, 220), (5, 300), (7, 33)]:
    dev = "cuda"
    torch.manual_seed(S)
    qkv = (torch.randn(B * S, 2304, device=dev)).bfloat16()
    lens = torch.randint(max(1, S // 2), S + 1, (B,), device=dev); lens[0] = S
    mask = (torch.arange(S, device=dev)[None, :] < lens[:, None]).long().contiguous()
    ctx = torch.empty(B * S, 768, dtype=torch.bfloat16, device=dev)
    lse = torch.empty(B, 12, S, device=dev)
    dctx = torch.randn(B * S, 768, device=dev).bfloat16()
    dqkv = torch.empty(B * S, 2304, dtype=torch.bfloat16, device=dev)
    delta = torch.empty(B, 12, S, device=dev)
    f = lambda: call("uc2_attention_fwd", qkv.data_ptr(), mask.data_ptr(), ctx.data_ptr(), lse.data_ptr(), B, S, stream())
    bw = lambda: call("uc2_attention_bwd", qkv.data_ptr(), mask.data_ptr(), ctx.data_ptr(), dctx.data_ptr(), lse.data_ptr(), delta.data_ptr(), dqkv.data_ptr(), B, S, stream())
    drop = (0x1234567, int(round(0.1 * 65536)), 1.0 / 0.9)
    fd = lambda: call("uc2_attention_fwd_dropout", qkv.data_ptr(), mask.data_ptr(), ctx.data_ptr(), lse.data_ptr(), B, S, *drop, stream())
    bd = lambda: call("uc2_attention_bwd_dropout", qkv.data_ptr(), mask.data_ptr(), ctx.data_ptr(), dctx.data_ptr(), lse.data_ptr(), delta.data_ptr(), dqkv.data_ptr(), B, S, *drop, stream())
    ufd = t(fd) if S <= 256 else float("nan"); ubd = t(bd) if S <= 256 else float("nan")
    uf = t(f); ub = t(bw)
    fl = 4.0 * B * 12 * S * S * 64
    # reference
    x = qkv.float().view(B, S, 3, 12, 64).permute(2, 0, 3, 1, 4).contiguous().requires_grad_(True)
    q, k, v = x[0], x[1], x[2]
    sc = q @ k.transpose(-1, -2) / 8 + (1 - mask.float())[:, None, None, :] * -10000.0
    p = sc.softmax(-1)
    o = (p @ v).permute(0, 2, 1, 3).reshape(B * S, 768)
    o.backward(dctx.float())
    gref = x.grad.permute(1, 3, 0, 2, 4).reshape(B * S, 2304)
    e_f = (ctx.float() - o).abs().max().item()
    e_b = (dqkv.float() - gref).abs().max().item() / gref.abs().max().item()
    print(f"B={B:4d} S={S:4d}  fwd {uf:7.1f} us {fl / uf / 1e6:6.1f} TFLOP/s   bwd {ub:7.1f} us {2.5 * fl / ub / 1e6:6.1f} TFLOP/s (5 products)   "
          f"max err fwd {e_f:.4f} bwd rel {e_b:.4f}   with dropout 0.1: fwd {ufd:7.1f} us bwd {ubd:7.1f} us")
