"""GPU: the tcgen05 GEMM (uc2_gemm_bf16) against an fp32 torch matmul of the same bf16 inputs.
Covers forward (K-major/K-major), dgrad (B MN-major), wgrad (both MN-major, split-K atomic
accumulate), every fused epilogue, ragged M/N/K edges and all tile widths."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _rand(shape, seed, scale=1.0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(shape, generator=g) * scale).cuda()


def _ref_gelu(x):
    return x * 0.5 * (1.0 + torch.erf(x / 2 ** 0.5))


def _ref_dgelu(x):
    return 0.5 * (1.0 + torch.erf(x / 2 ** 0.5)) + x * torch.exp(-0.5 * x * x) / (2 * torch.pi) ** 0.5


def _check(got, ref, tol, what):
    err = (got.float() - ref).abs().max().item()
    scale = ref.abs().max().item() + 1e-6
    assert err <= tol * scale, f"{what}: max err {err:.4g} vs scale {scale:.4g}"


@pytest.mark.parametrize("block_n", [0, 64, 128, 256])
@pytest.mark.parametrize("M,N,K", [(256, 768, 768), (19200 // 8, 2304, 768), (200, 3072, 768), (333, 768, 3072),
                                   (128, 64, 64), (77, 1601, 768)])
def test_forward_bias(M, N, K, block_n):
    from uc2_b200 import _lib
    a = _rand((M, K), 1).bfloat16()
    b = _rand((N, K), 2, 0.05).bfloat16()
    bias = _rand((N,), 3)
    out = torch.empty(M, N, dtype=torch.bfloat16, device="cuda")
    out32 = torch.empty(M, N, dtype=torch.float32, device="cuda")
    _lib.gemm(a, b, M, N, K, bias=bias, out_bf16=out, out_f32=out32, block_n=block_n)
    ref = a.float() @ b.float().t() + bias
    _check(out32, ref, 2e-5, "fp32 out")
    _check(out, ref, 1e-2, "bf16 out")


def test_epilogues():
    from uc2_b200 import _lib
    M, N, K = 384, 3072, 768
    a = _rand((M, K), 4).bfloat16()
    b = _rand((N, K), 5, 0.05).bfloat16()
    bias = _rand((N,), 6)
    res = _rand((M, N), 7).bfloat16()
    # bias + GELU with pre-activation copy
    g = torch.empty(M, N, dtype=torch.bfloat16, device="cuda")
    u = torch.empty(M, N, dtype=torch.bfloat16, device="cuda")
    _lib.gemm(a, b, M, N, K, bias=bias, act=_lib.ACT_GELU, out_bf16=g, out_pre=u)
    pre = a.float() @ b.float().t() + bias
    _check(u, pre, 1e-2, "pre-activation")
    _check(g, _ref_gelu(pre), 1e-2, "gelu")
    # bias + residual
    z = torch.empty(M, N, dtype=torch.bfloat16, device="cuda")
    _lib.gemm(a, b, M, N, K, bias=bias, residual=res, out_bf16=z)
    _check(z, pre + res.float(), 1e-2, "residual")
    # dGELU: acc * gelu'(aux)
    d = torch.empty(M, N, dtype=torch.bfloat16, device="cuda")
    _lib.gemm(a, b, M, N, K, aux=u, act=_lib.ACT_DGELU, out_bf16=d)
    _check(d, (a.float() @ b.float().t()) * _ref_dgelu(u.float()), 1e-2, "dgelu")
    # tanh
    t = torch.empty(M, N, dtype=torch.float32, device="cuda")
    _lib.gemm(a, b, M, N, K, bias=bias, act=_lib.ACT_TANH, out_f32=t)
    _check(t, torch.tanh(pre), 1e-4, "tanh")


@pytest.mark.parametrize("block_n", [64, 128, 256])
@pytest.mark.parametrize("M,N,K", [(256, 768, 3072), (1000, 3072, 768), (130, 768, 2304)])
def test_dgrad_b_mn_major(M, N, K, block_n):
    """dX[M,N] = dY[M,K] @ W[K,N]  (W stored [K][N]: MN-major B)."""
    from uc2_b200 import _lib
    dy = _rand((M, K), 8).bfloat16()
    w = _rand((K, N), 9, 0.05).bfloat16()
    res = _rand((M, N), 10).bfloat16()
    out = torch.empty(M, N, dtype=torch.bfloat16, device="cuda")
    _lib.gemm(dy, w, M, N, K, b_mn=True, residual=res, out_bf16=out, block_n=block_n)
    _check(out, dy.float() @ w.float() + res.float(), 1e-2, "dgrad")


@pytest.mark.parametrize("split_k", [1, 3, 8])
@pytest.mark.parametrize("block_n", [64, 128, 256])
@pytest.mark.parametrize("Mtok,N,K", [(1024, 768, 768), (2500, 3072, 768), (640, 768, 2048)])
def test_wgrad_both_mn_major(Mtok, N, K, block_n, split_k):
    """dW[N,K] += dY[Mtok,N]^T @ X[Mtok,K]: in GEMM terms M:=N, N:=K, contraction Mtok."""
    from uc2_b200 import _lib
    dy = _rand((Mtok, N), 11).bfloat16()
    x = _rand((Mtok, K), 12).bfloat16()
    dw = torch.ones(N, K, dtype=torch.float32, device="cuda")
    _lib.gemm(dy, x, N, K, Mtok, a_mn=True, b_mn=True, out_f32=dw, accumulate=True, split_k=split_k,
              block_n=block_n)
    ref = dy.float().t() @ x.float() + 1.0
    _check(dw, ref, 1e-4, "wgrad")


def test_a_mn_b_k():
    from uc2_b200 import _lib
    M, N, K = 256, 128, 320
    at = _rand((K, M), 13).bfloat16()
    b = _rand((N, K), 14).bfloat16()
    out = torch.empty(M, N, dtype=torch.float32, device="cuda")
    _lib.gemm(at, b, M, N, K, a_mn=True, out_f32=out)
    _check(out, at.float().t() @ b.float().t(), 2e-5, "A MN-major")


def test_strided_views():
    """Operands that are column slices of wider tensors (leading dimension > extent)."""
    from uc2_b200 import _lib
    M, N, K = 300, 768, 768
    big = _rand((M, 2304), 15).bfloat16()
    a = big[:, 768:1536]
    b = _rand((N, K), 16, 0.05).bfloat16()
    outbig = torch.zeros(M, 2304, dtype=torch.bfloat16, device="cuda")
    _lib.gemm(a, b, M, N, K, out_bf16=outbig[:, 1536:])
    _check(outbig[:, 1536:], a.float() @ b.float().t(), 1e-2, "strided")
    assert outbig[:, :1536].abs().max().item() == 0


def test_bad_args_raise():
    from uc2_b200 import _lib
    a = torch.zeros(128, 70, dtype=torch.bfloat16, device="cuda")
    b = torch.zeros(128, 70, dtype=torch.bfloat16, device="cuda")
    out = torch.zeros(128, 128, dtype=torch.bfloat16, device="cuda")
    with pytest.raises(RuntimeError):
        _lib.gemm(a, b, 128, 128, 70, out_bf16=out)     # pitch not a multiple of 8 elements


@pytest.mark.parametrize("ctas", [1, 2])
@pytest.mark.parametrize("block_n", [128, 256])
@pytest.mark.parametrize("M,N,K", [(512, 768, 768), (2400, 2304, 768), (300, 3072, 768), (1333, 768, 3072),
                                   (129, 1601, 768)])
def test_cta_pair_forward_and_epilogues(M, N, K, block_n, ctas):
    """CTA-pair (cta_group::2, 256-row tiles) and single-CTA forms of the same GEMM, every fused epilogue,
    ragged M (a pair whose second CTA is partly or wholly out of range) and ragged N."""
    from uc2_b200 import _lib
    a = _rand((M, K), 21).bfloat16()
    b = _rand((N, K), 22, 0.05).bfloat16()
    bias = _rand((N,), 23)
    Np = (N + 15) // 16 * 16
    res32 = _rand((M, Np), 24)[:, :N]
    resb = _rand((M, Np), 25).bfloat16()[:, :N]
    pre = a.float() @ b.float().t() + bias
    kw = dict(block_n=block_n, ctas=ctas)
    z = torch.empty(M, Np, dtype=torch.float32, device="cuda")[:, :N]
    _lib.gemm(a, b, M, N, K, bias=bias, residual=res32, out_f32=z, **kw)
    _check(z, pre + res32, 2e-5, "fp32 residual -> fp32")
    zb = torch.empty(M, Np, dtype=torch.bfloat16, device="cuda")[:, :N]
    _lib.gemm(a, b, M, N, K, bias=bias, residual=resb, out_bf16=zb, **kw)
    _check(zb, pre + resb.float(), 1e-2, "bf16 residual")
    g = torch.empty(M, Np, dtype=torch.bfloat16, device="cuda")[:, :N]
    u = torch.empty(M, Np, dtype=torch.bfloat16, device="cuda")[:, :N]
    _lib.gemm(a, b, M, N, K, bias=bias, act=_lib.ACT_GELU, out_bf16=g, out_pre=u, **kw)
    _check(u, pre, 1e-2, "pre-activation")
    _check(g, _ref_gelu(pre), 1e-2, "gelu")
    d = torch.empty(M, Np, dtype=torch.bfloat16, device="cuda")[:, :N]
    _lib.gemm(a, b, M, N, K, aux=u, act=_lib.ACT_DGELU, out_bf16=d, **kw)
    _check(d, (a.float() @ b.float().t()) * _ref_dgelu(u.float()), 1e-2, "dgelu")


@pytest.mark.parametrize("ctas", [1, 2])
@pytest.mark.parametrize("block_n", [128, 256])
def test_cta_pair_dgrad_wgrad(block_n, ctas):
    from uc2_b200 import _lib
    M, N, K = 1000, 768, 3072
    dy = _rand((M, K), 31).bfloat16()
    w = _rand((K, N), 32, 0.05).bfloat16()
    out = torch.empty(M, N, dtype=torch.bfloat16, device="cuda")
    _lib.gemm(dy, w, M, N, K, b_mn=True, out_bf16=out, block_n=block_n, ctas=ctas)
    _check(out, dy.float() @ w.float(), 1e-2, "dgrad")
    Mtok, N2, K2 = 2500, 3072, 768
    dy2 = _rand((Mtok, N2), 33).bfloat16()
    x = _rand((Mtok, K2), 34).bfloat16()
    for split in (0, 1, 5):
        dw = torch.ones(N2, K2, dtype=torch.float32, device="cuda")
        _lib.gemm(dy2, x, N2, K2, Mtok, a_mn=True, b_mn=True, out_f32=dw, accumulate=True, split_k=split,
                  block_n=block_n, ctas=ctas)
        _check(dw, dy2.float().t() @ x.float() + 1.0, 1e-4, f"wgrad split {split}")


@pytest.mark.parametrize("N,K", [(768, 768), (2304, 768)])
def test_bench_sized_token_dim_tail_split(N, K):
    """M = 19200 tokens (BASELINE configs[1]: 120 pairs x 160): 75 blocks of 256 rows on 74 CTA pairs -- the
    leftover block goes to a second small-tile launch; every row must still be written exactly once."""
    from uc2_b200 import _lib
    M = 19200
    a = _rand((M, K), 41).bfloat16()
    b = _rand((N, K), 42, 0.05).bfloat16()
    bias = _rand((N,), 43)
    res = _rand((M, N), 44)
    out = torch.full((M, N), float("nan"), dtype=torch.float32, device="cuda")
    ob = torch.empty((M, N), dtype=torch.bfloat16, device="cuda")
    _lib.gemm(a, b, M, N, K, bias=bias, residual=res, out_f32=out, out_bf16=ob)
    ref = a.float() @ b.float().t() + bias + res
    _check(out, ref, 2e-5, "fp32 out")
    _check(ob, ref, 1e-2, "bf16 out")
    assert torch.equal(out[-300:].bfloat16(), ob[-300:])


@pytest.mark.parametrize("held_sms", [0, 36, 100])
def test_dynamic_tile_order_with_sms_held_by_another_kernel(held_sms):
    """uc2_gemm_sched_dynamic: workers draw tiles from a device counter.  With SMs held by another kernel (what a NCCL
    collective of the overlapped gradient exchange does) some workers become resident only after the counter has run
    out: they must leave at once, every tile must still be computed exactly once, and the counter pair must be left
    zeroed for its next user.  Same bits as the fixed tile order (no split-K here)."""
    from uc2_b200 import _lib
    L = _lib.lib()
    M, N, K = 10240, 2304, 768
    a = _rand((M, K), 40).bfloat16()
    b = _rand((N, K), 41, 0.05).bfloat16()
    bias = _rand((N,), 42)
    aux = _rand((M, N), 43).bfloat16()
    ref = torch.empty(M, N, dtype=torch.bfloat16, device="cuda")
    ref_g = torch.empty_like(ref)
    dw_ref = torch.zeros(N, K, dtype=torch.float32, device="cuda")
    prev = L.uc2_gemm_sched_dynamic(0)
    try:
        _lib.gemm(a, b, M, N, K, bias=bias, out_bf16=ref)
        _lib.gemm(a, b, M, N, K, bias=bias, act=_lib.ACT_GELU, out_bf16=ref_g, out_pre=torch.empty_like(ref))
        _lib.gemm(ref, a, N, K, M, a_mn=True, b_mn=True, out_f32=dw_ref, accumulate=True, split_k=0)
        torch.cuda.synchronize()
        L.uc2_gemm_sched_dynamic(1)
        side = torch.cuda.Stream()
        outs = [torch.empty_like(ref) for _ in range(6)]
        outs_g = [torch.empty_like(ref) for _ in range(6)]
        pre = torch.empty_like(ref)
        dw = torch.zeros_like(dw_ref)
        if held_sms:
            # ~3 ms at 1.9 GHz: longer than the 13 GEMMs below, so late workers exist in every one of them
            assert L.uc2_debug_occupy_sms(held_sms, 6_000_000, side.cuda_stream) == 0
        for o, og in zip(outs, outs_g):
            _lib.gemm(a, b, M, N, K, bias=bias, out_bf16=o)
            _lib.gemm(a, b, M, N, K, bias=bias, act=_lib.ACT_GELU, out_bf16=og, out_pre=pre)
        _lib.gemm(ref, a, N, K, M, a_mn=True, b_mn=True, out_f32=dw, accumulate=True, split_k=0)
        torch.cuda.synchronize()
    finally:
        L.uc2_gemm_sched_dynamic(prev)
    for o, og in zip(outs, outs_g):
        assert torch.equal(o, ref)
        assert torch.equal(og, ref_g)
    _check(dw, dw_ref, 1e-5, "split-K wgrad under the dynamic tile order")
