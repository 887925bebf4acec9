"""GPU: the masked flash attention kernels (uc2_attention_fwd / _bwd) against a torch fp32 restatement of
BertSelfAttention (model/layer.py:80-100: scores / 8 + (1 - mask) * -10000, softmax, P V) on the same bf16 inputs.
Shapes cover the per-head kernels (S <= 256: double- and single-buffered backward, ragged S, one warp round and
two) and the tiled general path (S > 256), with prefix masks, arbitrary 0/1 masks and fully padded tails."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _ref(qkv, mask, dctx, B, S):
    x = qkv.float().view(B, S, 3, 12, 64).permute(2, 0, 3, 1, 4).contiguous().requires_grad_(True)
    q, k, v = x[0], x[1], x[2]
    sc = q @ k.transpose(-1, -2) / 8 + (1 - mask.float())[:, None, None, :] * -10000.0
    p = sc.softmax(-1)
    o = (p @ v).permute(0, 2, 1, 3).reshape(B * S, 768)
    o.backward(dctx.float())
    lse = torch.logsumexp(sc, -1)
    return o.detach(), x.grad.permute(1, 3, 0, 2, 4).reshape(B * S, 2304), lse


@pytest.mark.parametrize("B,S,kind", [(6, 160, "prefix"), (3, 76, "prefix"), (4, 220, "prefix"), (2, 256, "random"),
                                      (3, 300, "prefix"), (5, 33, "random"), (2, 16, "prefix"), (3, 208, "prefix"),
                                      (150, 160, "prefix")])
def test_attention_forward_backward(B, S, kind):
    from uc2_b200._lib import call, stream
    torch.manual_seed(B * 1000 + S)
    dev = "cuda"
    qkv = torch.randn(B * S, 2304, device=dev).bfloat16()
    if kind == "prefix":
        lens = torch.randint(max(1, S // 3), S + 1, (B,), device=dev)
        lens[0] = S
        mask = (torch.arange(S, device=dev)[None, :] < lens[:, None]).long().contiguous()
    else:
        mask = (torch.rand(B, S, device=dev) < 0.7).long()
        mask[:, 0] = 1
    ctx = torch.full((B * S, 768), float("nan"), dtype=torch.bfloat16, device=dev)
    lse = torch.empty(B, 12, S, device=dev)
    dctx = torch.randn(B * S, 768, device=dev).bfloat16()
    dqkv = torch.full((B * S, 2304), float("nan"), dtype=torch.bfloat16, device=dev)
    delta = torch.empty(B, 12, S, device=dev)
    call("uc2_attention_fwd", qkv.data_ptr(), mask.data_ptr(), ctx.data_ptr(), lse.data_ptr(), B, S, stream())
    call("uc2_attention_bwd", qkv.data_ptr(), mask.data_ptr(), ctx.data_ptr(), dctx.data_ptr(), lse.data_ptr(),
         delta.data_ptr(), dqkv.data_ptr(), B, S, stream())
    o, g, lse_ref = _ref(qkv, mask, dctx, B, S)
    assert torch.isfinite(ctx.float()).all() and torch.isfinite(dqkv.float()).all()
    assert (ctx.float() - o).abs().max().item() <= 2e-2
    assert (lse - lse_ref).abs().max().item() <= 2e-3
    scale = g.abs().max().item()
    assert (dqkv.float() - g).abs().max().item() <= 1.5e-2 * scale
