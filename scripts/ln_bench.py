#!/usr/bin/env python
"""LayerNorm fwd/bwd and column-sum bandwidth at the bench token count (HBM roofline check)."""
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

M, H = int(sys.argv[1]) if len(sys.argv) > 1 else 19200, 768
dev = "cuda"
x32 = torch.randn(M, H, device=dev); dy = torch.randn(M, H, device=dev).bfloat16()
g = torch.ones(H, device=dev); b = torch.zeros(H, device=dev)
y = torch.empty(M, H, dtype=torch.bfloat16, device=dev); y32 = torch.empty(M, H, device=dev)
dx = torch.empty(M, H, dtype=torch.bfloat16, device=dev)
dg, db, dbias = (torch.zeros(H, device=dev) for _ in range(3))
# several distinct buffers so consecutive launches do not hit L2
xs = [torch.randn(M, H, device=dev) for _ in range(4)]
i = [0]
def fwd():
    i[0] += 1
    call("uc2_layernorm_fwd", xs[i[0] % 4].data_ptr(), 1, g.data_ptr(), b.data_ptr(), 1e-12, y.data_ptr(), y32.data_ptr(), M, stream())
def bwd():
    i[0] += 1
    call("uc2_layernorm_bwd", xs[i[0] % 4].data_ptr(), 1, dy.data_ptr(), g.data_ptr(), 1e-12, dx.data_ptr(), dg.data_ptr(), db.data_ptr(), dbias.data_ptr(), M, stream())
us = t(fwd); print(f"layernorm_fwd  {us:7.1f} us  {(M*H*(4+2+4))/us/1e3:7.1f} GB/s")
us = t(bwd); print(f"layernorm_bwd  {us:7.1f} us  {(M*H*(4+2+2))/us/1e3:7.1f} GB/s")
big = torch.randn(M, 3072, device=dev).bfloat16(); out = torch.zeros(3072, device=dev)
us = t(lambda: call("uc2_colsum_bf16", big.data_ptr(), 3072, M, 3072, out.data_ptr(), stream()))
print(f"colsum [M,3072] {us:7.1f} us  {(M*3072*2)/us/1e3:7.1f} GB/s")
# correctness of bwd vs torch
xr = xs[0].clone().requires_grad_(True); gr = torch.randn(H, device=dev).requires_grad_(True)
yr = torch.nn.functional.layer_norm(xr, (H,), gr, b, 1e-12); yr.backward(dy.float())
dg.zero_(); db.zero_(); dbias.zero_()
call("uc2_layernorm_bwd", xs[0].data_ptr(), 1, dy.data_ptr(), gr.data_ptr(), 1e-12, dx.data_ptr(), dg.data_ptr(), db.data_ptr(), dbias.data_ptr(), M, stream())
torch.cuda.synchronize()
print("dx err", (dx.float() - xr.grad).abs().max().item(), "dgamma rel", ((dg - gr.grad).norm() / gr.grad.norm()).item(),
      "dbias rel", ((dbias - xr.grad.sum(0)).norm() / xr.grad.sum(0).norm()).item())
