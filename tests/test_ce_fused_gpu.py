"""GPU: cross entropy fused into the decoder GEMM (csrc/gemm_tcgen05.cu MODE 9 + uc2_ce_stats_reduce +
uc2_ce_bwd_inplace_bf16) against torch on the same bf16 operands -- the arithmetic of RobertaLMHead's decoder
(model/layer.py:263-264) followed by F.cross_entropy(reduction='none') (model/model.py:592-596) and its backward.
Shapes cover the CTA-pair and single-CTA tilings, a vocabulary that is not a multiple of 32 (ragged last chunk, the
XLM-R case: 250 002), a multiple of 256, labels in the first / last column, and ignored rows."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _run(n, N, K, seed, ignore=()):
    from uc2_b200 import _lib
    from uc2_b200._lib import call, stream
    g = torch.Generator(device="cuda").manual_seed(seed)
    h = (torch.randn(n, K, device="cuda", generator=g) * 0.5).bfloat16()
    W = (torch.randn(N, K, device="cuda", generator=g) * 0.1).bfloat16()
    bias = torch.randn(N, device="cuda", generator=g) * 0.3
    t = torch.randint(0, N, (n,), device="cuda", generator=g)
    t[0], t[-1] = 0, N - 1
    for i in ignore:
        t[i] = -1
    pitch = (N + 15) // 16 * 16
    logits = torch.full((n, pitch), float("nan"), dtype=torch.bfloat16, device="cuda")
    n_chunks = (N + 31) // 32
    stats = torch.full((n_chunks, n, 2), float("nan"), device="cuda")
    part = torch.empty(((n_chunks + 255) // 256, n, 2), device="cuda")
    tgt = torch.zeros(n, device="cuda")
    loss = torch.empty(n, device="cuda")
    lse = torch.empty(n, device="cuda")
    _lib.gemm(h, W, n, N, K, bias=bias, out_bf16=logits[:, :N], ld_out=pitch, ce=(stats, t, tgt))
    call("uc2_ce_stats_reduce", stats.data_ptr(), n, n_chunks, n, tgt.data_ptr(), t.data_ptr(), -1, part.data_ptr(),
         loss.data_ptr(), lse.data_ptr(), stream())
    torch.cuda.synchronize()
    z = h.float() @ W.float().t() + bias
    return h, W, bias, t, logits, loss, lse, z


@pytest.mark.parametrize("n,N,K", [(300, 250002, 768), (64, 8192, 768), (129, 1000, 64), (5, 33, 128), (700, 4099, 768)])
def test_fused_ce_forward_matches_torch(n, N, K):
    h, W, bias, t, logits, loss, lse, z = _run(n, N, K, seed=n + N, ignore=(1,) if n > 2 else ())
    assert torch.isfinite(logits[:, :N].float()).all()
    assert (logits[:, :N].float() - z).abs().max().item() <= 2e-2 * max(1.0, z.abs().max().item() / 4)   # bf16 store
    ref_lse = torch.logsumexp(z, -1)
    assert (lse - ref_lse).abs().max().item() <= 2e-4 * ref_lse.abs().max().item() + 1e-4
    ref = torch.nn.functional.cross_entropy(z, t, ignore_index=-1, reduction="none")
    assert (loss - ref).abs().max().item() <= 1e-3
    assert float(loss[1]) == 0.0 or n <= 2


@pytest.mark.parametrize("n,N,K", [(300, 250002, 768), (129, 1000, 64), (5, 33, 128)])
def test_fused_ce_backward_in_place(n, N, K):
    from uc2_b200._lib import call, stream
    h, W, bias, t, logits, loss, lse, z = _run(n, N, K, seed=7 * n + N, ignore=(2,) if n > 3 else ())
    dloss = torch.rand(n, device="cuda") + 0.5
    zin = logits[:, :N].float().clone()
    call("uc2_ce_bwd_inplace_bf16", logits.data_ptr(), logits.stride(0), n, N, t.data_ptr(), -1, dloss.data_ptr(),
         lse.data_ptr(), stream())
    torch.cuda.synchronize()
    onehot = torch.zeros_like(zin)
    valid = t >= 0
    onehot[valid.nonzero().squeeze(1), t[valid]] = 1.0
    want = (torch.exp(zin - lse[:, None]) - onehot) * (dloss * valid)[:, None]
    got = logits[:, :N].float()
    assert (got - want).abs().max().item() <= 8e-3 * want.abs().max().item() + 1e-6     # bf16 rounding of the result
    # and against the exact gradient of the fp32 problem
    exact = (torch.softmax(z, -1) - onehot) * (dloss * valid)[:, None]
    assert (got - exact).abs().max().item() <= 3e-2 * exact.abs().max().item()


def test_fused_ce_rejects_other_epilogues():
    from uc2_b200 import _lib
    h = torch.zeros(8, 64, dtype=torch.bfloat16, device="cuda")
    W = torch.zeros(64, 64, dtype=torch.bfloat16, device="cuda")
    out = torch.empty(8, 64, dtype=torch.bfloat16, device="cuda")
    stats = torch.empty(2, 8, 2, device="cuda")
    t = torch.zeros(8, dtype=torch.long, device="cuda")
    tgt = torch.zeros(8, device="cuda")
    with pytest.raises(RuntimeError):
        _lib.gemm(h, W, 8, 64, 64, out_bf16=out, ce=(stats, t, tgt))          # no bias
