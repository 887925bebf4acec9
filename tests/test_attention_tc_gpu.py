"""GPU: the tcgen05 / TMEM attention kernels (csrc/attention_tc.cu: uc2_attention_fwd_tc, uc2_attention_bwd_tc)
against (1) a torch fp32 restatement of BertSelfAttention (model/layer.py:80-100) on the same bf16 inputs and (2) the
mma.sync kernels of csrc/attention.cu, with and without attention-probability dropout (same counter-hash stream,
so the kept set is identical and the outputs agree to bf16 rounding of P)."""
import contextlib

import pytest
import torch

pytestmark = [pytest.mark.gpu]


@contextlib.contextmanager
def _tc(on):
    """Route the public attention entry points to the tcgen05 kernels (1) or to the mma.sync ones (0)."""
    from uc2_b200._lib import lib
    prev = lib().uc2_attention_tc_enable(on)
    try:
        yield
    finally:
        lib().uc2_attention_tc_enable(prev)


SHAPES = [(6, 160, "prefix"), (3, 76, "prefix"), (2, 128, "random"), (5, 33, "random"), (2, 16, "prefix"),
          (3, 150, "prefix"), (4, 129, "random"), (150, 160, "prefix")]
# packed lengths up to 256 (BASELINE configs[4], VTLM: S = 222): one CTA per SM in the forward, two query blocks in the backward
FWD_SHAPES = SHAPES + [(3, 161, "random"), (4, 222, "prefix"), (2, 192, "random"), (3, 256, "prefix"), (40, 222, "prefix")]


def _inputs(B, S, kind):
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
    return qkv, mask


def _ref(qkv, mask, B, S):
    x = qkv.float().view(B, S, 3, 12, 64).permute(2, 0, 3, 1, 4)
    sc = x[0] @ x[1].transpose(-1, -2) / 8 + (1 - mask.float())[:, None, None, :] * -10000.0
    o = (sc.softmax(-1) @ x[2]).permute(0, 2, 1, 3).reshape(B * S, 768)
    return o, torch.logsumexp(sc, -1)


def _run(name, qkv, mask, B, S, drop):
    from uc2_b200._lib import call, stream
    if not name.endswith("_tc"):
        with _tc(0):                              # the public entry points, pinned to the mma.sync kernels
            return _run_raw(name, qkv, mask, B, S, drop)
    return _run_raw(name, qkv, mask, B, S, drop)


def _run_raw(name, qkv, mask, B, S, drop):
    from uc2_b200._lib import call, stream
    ctx = torch.full((B * S, 768), float("nan"), dtype=torch.bfloat16, device="cuda")
    lse = torch.full((B, 12, S), float("nan"), device="cuda")
    call(name, qkv.data_ptr(), mask.data_ptr(), ctx.data_ptr(), lse.data_ptr(), B, S, *drop, stream())
    torch.cuda.synchronize()
    return ctx, lse


@pytest.mark.parametrize("B,S,kind", FWD_SHAPES)
def test_tc_forward_matches_reference_and_mma_sync(B, S, kind):
    qkv, mask = _inputs(B, S, kind)
    ctx, lse = _run("uc2_attention_fwd_tc", qkv, mask, B, S, (0, 0, 1.0))
    o, lse_ref = _ref(qkv, mask, B, S)
    assert torch.isfinite(ctx.float()).all() and torch.isfinite(lse).all()
    assert (ctx.float() - o).abs().max().item() <= 2e-2          # north_star: hidden states 2e-2 abs in bf16
    assert (lse - lse_ref).abs().max().item() <= 2e-3
    ctx0, lse0 = _run("uc2_attention_fwd_dropout", qkv, mask, B, S, (0, 0, 1.0))
    assert (ctx.float() - ctx0.float()).abs().max().item() <= 2e-2
    assert (lse - lse0).abs().max().item() <= 1e-4


def _keep_masks(key, B, S, p):
    """[B, 12, S, S] keep masks of the tcgen05 kernels' attention dropout stream (host mirror, uc2_b200/dropout.py)."""
    import numpy as np
    from uc2_b200 import dropout as DO
    t = DO.thresh_of(p)
    m = np.stack([DO.attn_keep_mask_np(DO.head_key(key, bh), S, t) for bh in range(B * 12)]).reshape(B, 12, S, S)
    return torch.from_numpy(m).cuda(), t, DO.scale_of(p)


@pytest.mark.parametrize("B,S,kind", [(6, 160, "prefix"), (3, 77, "random"), (4, 150, "prefix"), (3, 222, "prefix"),
                                      (2, 256, "random"), (5, 33, "prefix")])
def test_tc_forward_dropout_matches_reference_with_host_mirror_mask(B, S, kind):
    """layer.py:94: dropout on the probabilities after the softmax (the normaliser keeps every key).  The kernel
    regenerates the mask from (key, query, key index); the torch reference gets the same mask from the host mirror."""
    qkv, mask = _inputs(B, S, kind)
    key, p = 0x1234567, 0.1
    keep, t, scale = _keep_masks(key, B, S, p)
    ctx, lse = _run("uc2_attention_fwd_tc", qkv, mask, B, S, (key, t, scale))
    x = qkv.float().view(B, S, 3, 12, 64).permute(2, 0, 3, 1, 4)
    sc = x[0] @ x[1].transpose(-1, -2) / 8 + (1 - mask.float())[:, None, None, :] * -10000.0
    o = ((sc.softmax(-1) * keep * scale) @ x[2]).permute(0, 2, 1, 3).reshape(B * S, 768)
    assert abs(keep.float().mean().item() - (1 - p)) < 5e-3
    assert torch.isfinite(ctx.float()).all()
    # a single wrong keep bit moves an output by O(p_max * |v|) ~ 0.1-1, far above this bar
    assert (ctx.float() - o).abs().max().item() <= 2e-2
    assert (lse - torch.logsumexp(sc, -1)).abs().max().item() <= 2e-3


def test_tc_switch_routes_the_public_entry_point():
    from uc2_b200._lib import lib
    B, S = 4, 160
    qkv, mask = _inputs(B, S, "prefix")
    with _tc(1):
        ctx, lse = _run_raw("uc2_attention_fwd_dropout", qkv, mask, B, S, (0, 0, 1.0))
    ctx_tc, lse_tc = _run("uc2_attention_fwd_tc", qkv, mask, B, S, (0, 0, 1.0))
    assert torch.equal(ctx, ctx_tc) and torch.equal(lse, lse_tc)


def test_tc_rejects_long_sequences():
    from uc2_b200._lib import lib, stream
    B, S = 1, 272
    qkv, mask = _inputs(B, S, "prefix")
    ctx = torch.empty(B * S, 768, dtype=torch.bfloat16, device="cuda")
    lse = torch.empty(B, 12, S, device="cuda")
    rc = lib().uc2_attention_fwd_tc(qkv.data_ptr(), mask.data_ptr(), ctx.data_ptr(), lse.data_ptr(), B, S, 0, 0, 1.0,
                                    stream())
    assert rc != 0


def _ref_bwd(qkv, mask, dctx, B, S, keep=None, scale=1.0):
    x = qkv.float().view(B, S, 3, 12, 64).permute(2, 0, 3, 1, 4).contiguous().requires_grad_(True)
    sc = x[0] @ x[1].transpose(-1, -2) / 8 + (1 - mask.float())[:, None, None, :] * -10000.0
    pr = sc.softmax(-1)
    if keep is not None:
        pr = pr * keep * scale
    o = (pr @ x[2]).permute(0, 2, 1, 3).reshape(B * S, 768)
    o.backward(dctx.float())
    return x.grad.permute(1, 3, 0, 2, 4).reshape(B * S, 2304)


def _run_bwd(name, qkv, mask, ctx, dctx, lse, B, S, drop):
    if not name.endswith("_tc"):
        with _tc(0):
            return _run_bwd_raw(name, qkv, mask, ctx, dctx, lse, B, S, drop)
    return _run_bwd_raw(name, qkv, mask, ctx, dctx, lse, B, S, drop)


def _run_bwd_raw(name, qkv, mask, ctx, dctx, lse, B, S, drop):
    from uc2_b200._lib import call, stream
    dqkv = torch.full((B * S, 2304), float("nan"), dtype=torch.bfloat16, device="cuda")
    if name == "uc2_attention_bwd_tc":
        call(name, qkv.data_ptr(), mask.data_ptr(), ctx.data_ptr(), dctx.data_ptr(), lse.data_ptr(), dqkv.data_ptr(),
             B, S, *drop, stream())
    else:
        delta = torch.empty(B, 12, S, device="cuda")
        call(name, qkv.data_ptr(), mask.data_ptr(), ctx.data_ptr(), dctx.data_ptr(), lse.data_ptr(), delta.data_ptr(),
             dqkv.data_ptr(), B, S, *drop, stream())
    torch.cuda.synchronize()
    return dqkv


@pytest.mark.parametrize("B,S,kind", FWD_SHAPES)
def test_tc_backward_matches_reference_and_mma_sync(B, S, kind):
    qkv, mask = _inputs(B, S, kind)
    dctx = torch.randn(B * S, 768, device="cuda").bfloat16()
    ctx, lse = _run("uc2_attention_fwd_dropout", qkv, mask, B, S, (0, 0, 1.0))
    dqkv = _run_bwd("uc2_attention_bwd_tc", qkv, mask, ctx, dctx, lse, B, S, (0, 0, 1.0))
    g = _ref_bwd(qkv, mask, dctx, B, S)
    assert torch.isfinite(dqkv.float()).all()
    scale = g.abs().max().item()
    assert (dqkv.float() - g).abs().max().item() <= 1.5e-2 * scale     # the bar of tests/test_attention_gpu.py
    d0 = _run_bwd("uc2_attention_bwd_dropout", qkv, mask, ctx, dctx, lse, B, S, (0, 0, 1.0))
    assert (dqkv.float() - d0.float()).abs().max().item() <= 1.5e-2 * scale


@pytest.mark.parametrize("B,S,kind", [(6, 160, "prefix"), (3, 77, "random"), (4, 150, "prefix"), (5, 33, "prefix"),
                                      (3, 222, "prefix"), (2, 256, "random"), (2, 176, "random")])
def test_tc_backward_dropout_matches_reference_with_host_mirror_mask(B, S, kind):
    """Forward and backward regenerate the same attention-dropout mask: torch autograd through the dropped
    probabilities (mask from the host mirror) gives the gradients the kernel must produce."""
    qkv, mask = _inputs(B, S, kind)
    dctx = torch.randn(B * S, 768, device="cuda").bfloat16()
    key, p = 0x7654321, 0.1
    keep, t, scale = _keep_masks(key, B, S, p)
    ctx, lse = _run("uc2_attention_fwd_tc", qkv, mask, B, S, (key, t, scale))
    d_tc = _run_bwd("uc2_attention_bwd_tc", qkv, mask, ctx, dctx, lse, B, S, (key, t, scale))
    g = _ref_bwd(qkv, mask, dctx, B, S, keep, scale)
    sc = g.abs().max().item()
    assert torch.isfinite(d_tc.float()).all()
    assert (d_tc.float() - g).abs().max().item() <= 1.5e-2 * sc


@pytest.mark.parametrize("on", [0, 1])
def test_training_step_end_to_end_vs_oracle(on):
    """A small ITM rank model (2 layers) forward + backward with the mma.sync (0) / tcgen05 (1) attention kernels
    against oracle autograd (fp32 CPU): same loss, every parameter gradient within the relative L2 bar of
    tests/test_model_gpu.py::test_full_gradients_vs_oracle for the ITM losses (a difference of near-equal pooled
    vectors on a random-init network amplifies the bf16 forward error)."""
    import cases
    from oracle import uc2_oracle as O
    from test_model_gpu import build, dev
    from uc2_b200.utils import set_dropout
    cfg = cases.config(layers=2)
    b = cases.batch_rank()
    m, sd = build("retrieval", cfg)
    m.train()
    set_dropout(m, 0)
    sdg = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    lref = O.forward_retrieval(sdg, O.Family("vlxlmr"), b).mean()
    lref.backward()
    with _tc(on):
        loss = m(dev(b), compute_loss=True).mean()
        loss.backward()
        torch.cuda.synchronize()
    assert abs(loss.item() - lref.item()) <= 2e-3 * max(abs(lref.item()), 1e-6)
    top = max(float(v.grad.norm()) for v in sdg.values() if v.grad is not None)
    for n_, p_ in m.named_parameters():
        gr = sdg[n_].grad
        gg = p_.grad.detach().cpu().float()
        if gr is None or float(gr.norm()) < 1e-6 * top:
            assert float(gg.norm()) < 1e-3 * top, n_
            continue
        rel = float((gg - gr).norm() / gr.norm())
        # the triplet loss on a random-init network is a difference of near-equal sigmoid scores: it amplifies the bf16
        # forward rounding (both attention paths sit at 9-15% on the embedding-table gradients); a wrong kernel is O(1) off
        assert rel <= 2.5e-1, f"{n_}: relative gradient error {rel:.4f} (tc={on})"
