"""GPU: LayerNorm forward / backward (fp32 and bf16 inputs, ragged row counts) and the column sum, through the
C-ABI, against a plain torch fp32 layer_norm (FusedLayerNorm call sites of model/layer.py:108,149,196,242).
The dropout form also returns the masked, rescaled gradient of the dense branch; its mask is checked against the
host mirror of the counter hash (uc2_b200/dropout.py)."""
import pytest
import torch

pytestmark = pytest.mark.gpu

H = 768


def _ref(x, dy, gamma, beta, eps):
    xr = x.float().clone().requires_grad_(True)
    gr = gamma.clone().requires_grad_(True)
    br = beta.clone().requires_grad_(True)
    y = torch.nn.functional.layer_norm(xr, (H,), gr, br, eps)
    y.backward(dy.float())
    return y.detach(), xr.grad, gr.grad, br.grad


@pytest.mark.parametrize("rows", [1, 7, 130, 1187, 19203])
@pytest.mark.parametrize("f32", [True, False])
def test_layernorm_fwd_bwd(rows, f32):
    from uc2_b200._lib import call, stream
    g = torch.Generator(device="cuda").manual_seed(rows)
    x = torch.randn(rows, H, device="cuda", generator=g) * 1.7 + 0.4
    if not f32:
        x = x.bfloat16()
    dy = torch.randn(rows, H, device="cuda", generator=g).bfloat16()
    gamma = 1 + 0.1 * torch.randn(H, device="cuda", generator=g)
    beta = 0.1 * torch.randn(H, device="cuda", generator=g)
    eps = 1e-12
    y = torch.empty(rows, H, dtype=torch.bfloat16, device="cuda")
    y32 = torch.empty(rows, H, device="cuda")
    call("uc2_layernorm_fwd", x.data_ptr(), int(f32), gamma.data_ptr(), beta.data_ptr(), eps, y.data_ptr(),
         y32.data_ptr(), rows, stream())
    dx = torch.empty(rows, H, dtype=torch.bfloat16, device="cuda")
    dg, db, dbias = (torch.zeros(H, device="cuda") for _ in range(3))
    call("uc2_layernorm_bwd", x.data_ptr(), int(f32), dy.data_ptr(), gamma.data_ptr(), eps, dx.data_ptr(),
         dg.data_ptr(), db.data_ptr(), dbias.data_ptr(), rows, stream())
    ry, rdx, rdg, rdb = _ref(x, dy, gamma, beta, eps)
    torch.testing.assert_close(y32, ry, rtol=1e-5, atol=2e-5)
    torch.testing.assert_close(y.float(), ry, rtol=8e-3, atol=8e-3)
    torch.testing.assert_close(dx.float(), rdx, rtol=8e-3, atol=8e-3)
    scale = max(1.0, float(rows) ** 0.5)
    torch.testing.assert_close(dg, rdg, rtol=1e-3, atol=1e-3 * scale)
    torch.testing.assert_close(db, rdb, rtol=1e-3, atol=1e-3 * scale)
    # dbias = column sum of the bf16-rounded dx that was written
    torch.testing.assert_close(dbias, rdx.sum(0), rtol=2e-2, atol=2e-2 * scale)


@pytest.mark.parametrize("rows", [5, 2050])
def test_layernorm_bwd_dropout_mask(rows):
    from uc2_b200 import dropout as DR
    from uc2_b200._lib import call, stream
    g = torch.Generator(device="cuda").manual_seed(3)
    x = torch.randn(rows, H, device="cuda", generator=g)
    dy = torch.randn(rows, H, device="cuda", generator=g).bfloat16()
    gamma = 1 + 0.1 * torch.randn(H, device="cuda", generator=g)
    key, p = 0x1234ABCD, 0.1
    thresh, scale = DR.thresh_of(p), DR.scale_of(p)
    dx = torch.empty(rows, H, dtype=torch.bfloat16, device="cuda")
    dxm = torch.empty_like(dx)
    dg, db, dbias = (torch.zeros(H, device="cuda") for _ in range(3))
    call("uc2_layernorm_bwd_dropout", x.data_ptr(), 1, dy.data_ptr(), gamma.data_ptr(), 1e-12, dx.data_ptr(),
         dg.data_ptr(), db.data_ptr(), dbias.data_ptr(), rows, dxm.data_ptr(), key, thresh, scale, stream())
    _, rdx, _, _ = _ref(x, dy, gamma, torch.zeros(H, device="cuda"), 1e-12)
    torch.testing.assert_close(dx.float(), rdx, rtol=8e-3, atol=8e-3)
    keep = torch.from_numpy(DR.keep_mask_np(key, rows * H, thresh).reshape(rows, H)).cuda()
    want = torch.where(keep, rdx * scale, torch.zeros_like(rdx))
    torch.testing.assert_close(dxm.float(), want, rtol=8e-3, atol=8e-3)
    assert bool(((dxm == 0) | keep).all())                        # dropped elements are exact zeros
    assert 0.08 < 1 - float(keep.float().mean()) < 0.12
    # dbias accumulates the MASKED gradient (it belongs to the dense branch)
    torch.testing.assert_close(dbias, want.sum(0), rtol=2e-2, atol=2e-2 * max(1.0, rows ** 0.5))


def test_colsum_bf16():
    from uc2_b200._lib import call, stream
    g = torch.Generator(device="cuda").manual_seed(9)
    for rows, cols in ((3, 768), (4099, 3072), (257, 1608)):
        x = torch.randn(rows, cols, device="cuda", generator=g).bfloat16()
        out = torch.zeros(cols, device="cuda")
        call("uc2_colsum_bf16", x.data_ptr(), cols, rows, cols, out.data_ptr(), stream())
        torch.testing.assert_close(out, x.float().sum(0), rtol=1e-4, atol=1e-3)
