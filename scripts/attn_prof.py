#!/usr/bin/env python
"""Debug: UC2_ATTN_PROF=1 python scripts/attn_prof.py -- per-warp phase cycle counters of the tcgen05 attention forward."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from uc2_b200._lib import call, stream
B, S = 120, 160
qkv = torch.randn(B * S, 2304, device="cuda").bfloat16()
mask = torch.ones(B, S, dtype=torch.long, device="cuda")
ctx = torch.empty(B * S, 768, dtype=torch.bfloat16, device="cuda")
lse = torch.empty(B, 12, S, device="cuda")
for drop in [(0, 0, 1.0), (0x1234567, 6554, 1 / 0.9)]:
    for _ in range(2):
        call("uc2_attention_fwd_tc", qkv.data_ptr(), mask.data_ptr(), ctx.data_ptr(), lse.data_ptr(), B, S, *drop, stream())
dctx = torch.randn(B * S, 768, device="cuda").bfloat16()
dqkv = torch.empty(B * S, 2304, dtype=torch.bfloat16, device="cuda")
for drop in [(0, 0, 1.0), (0x1234567, 6554, 1 / 0.9)]:
    call("uc2_attention_fwd_tc", qkv.data_ptr(), mask.data_ptr(), ctx.data_ptr(), lse.data_ptr(), B, S, *drop, stream())
    for _ in range(2):
        call("uc2_attention_bwd_tc", qkv.data_ptr(), mask.data_ptr(), ctx.data_ptr(), dctx.data_ptr(), lse.data_ptr(),
             dqkv.data_ptr(), B, S, *drop, stream())
torch.cuda.synchronize()
