"""CPU walk-through of the DATA PATH of csrc/attention_tc.cu (the experimental tcgen05 attention) on real numbers.

Third leg of what can be checked without a GPU (beside the operand-layout model and the barrier-protocol
simulation): one (batch, head) item is pushed through a software model of the machine -- shared memory filled by
TMA boxes, tcgen05.mma as `D[128 x N] (+)= A B^T` over operands gathered through the descriptors, TMEM as a
128-lane x 512-column array read 32 lanes at a time, the element-wise code of the softmax / backward warps with
their lane-to-row mapping, swizzled 16-byte stores, the epilogues' row / column / output-slot mapping -- in the
order the kernels' barriers enforce, and the resulting ctx, lse, dQ, dK, dV are compared with a float64
restatement of BertSelfAttention (model/layer.py:80-100) and its gradient under the same dropout mask.  Rows past
S inside the TMA boxes carry the NEXT sample's (random) data and never-written shared memory is NaN, so a missing
mask, a wrong padding assumption or a swapped output slot shows up as a wrong or non-finite result.
The walk-through mirrors the kernels' constants and index arithmetic statement by statement."""
import numpy as np
import pytest
import torch

from test_attention_tc_layout_cpu import P_CHUNK, Smem
from uc2_b200 import dropout as DR

LOG2E = 1.4426950408889634
SCALE_LOG2 = np.float32(0.125 * LOG2E)
MASK_LOG2 = np.float32(-10000.0 * LOG2E)
HD = 64


def bf(x):
    return torch.as_tensor(np.asarray(x, np.float32)).bfloat16().float().numpy().astype(np.float64)


class Tmem:
    def __init__(self):
        self.v = np.full((128, 512), np.nan, np.float64)

    def umma(self, col, n, a, b, accumulate):
        d = a @ b.T                                      # a [128 x 16], b [n x 16]; NaN rows of a stay in their rows
        assert np.isfinite(b).all(), "a B operand (shared by every row) must never contain padding garbage"
        self.v[:, col:col + n] = (self.v[:, col:col + n] if accumulate else 0.0) + d

    def ld(self, q, col, n):                             # tcgen05.ld 32 lanes of quarter q, n columns, as fp32
        return self.v[32 * q:32 * q + 32, col:col + n].astype(np.float32)


def active_warps(S, t):
    rows = S - 128 * t
    return 0 if rows <= 0 else (4 if rows >= 128 else (rows + 31) // 32)


def make_case(S, seed, p_drop):
    rng = np.random.default_rng(seed)
    SP = (S + 15) // 16 * 16
    box = lambda: bf(rng.standard_normal((SP, HD)))      # rows >= S: the next sample's rows, as TMA delivers them
    Q, K, V, dO = box(), box(), box(), box()
    mask = (rng.random(S) < 0.75).astype(np.int64)
    mask[0] = 1
    thresh = DR.thresh_of(p_drop)
    hkey = DR.head_key(0x1234567, 5)
    qi, ki = np.meshgrid(np.arange(S), np.arange(S), indexing="ij")
    keep = DR.keep_mask_np(hkey, None, thresh, qi * S + ki) if thresh else np.ones((S, S), bool)
    return SP, Q, K, V, dO, mask, thresh, np.float32(DR.scale_of(p_drop)), hkey, keep


def reference(S, Q, K, V, dO, mask, keep, scale):
    q, k, v, do = Q[:S], K[:S], V[:S], dO[:S]
    sc = q @ k.T / 8 + (1 - mask)[None, :] * -10000.0
    mx = sc.max(1, keepdims=True)
    e = np.exp(sc - mx)
    P = e / e.sum(1, keepdims=True)
    lse = np.log(e.sum(1)) + mx[:, 0]
    Pd = P * keep * float(scale)
    O = Pd @ v
    dV = Pd.T @ do
    dP = (do @ v.T) * keep * float(scale)
    delta = (do * bf(O)).sum(1)
    dS = P * (dP - delta[:, None])
    return O, lse, dS @ k / 8, dS.T @ q / 8, dV


def forward_walk(S, SP, Q, K, V, mask, thresh, scale, hkey):
    nt, nchunk = (2 if S > 128 else 1), (SP + 63) // 64
    tile = SP * 128
    off_p = 2 * 3 * tile
    sm, tm = Smem(off_p + nt * nchunk * P_CHUNK + 2 * SP * 4 + 128), Tmem()
    TM_S, TM_S_STRIDE, TM_O, TM_O_STRIDE = 0, 192, 384, 64
    sQ, sK, sV = 0, tile, 2 * tile                                   # buffer 0
    sm.tma_box(sQ, Q); sm.tma_box(sK, K); sm.tma_box(sV, V)
    mb = np.where(np.arange(SP) < S, np.where(np.pad(mask, (0, SP - S)) != 0, np.float32(0), MASK_LOG2),
                  np.float32(-np.inf)).astype(np.float32)
    for t in range(nt):                                              # MMA warp: S_t = Q_t K^T
        for k in range(4):
            tm.umma(TM_S + t * TM_S_STRIDE, SP, sm.read_k_major(sQ + t * 16384 + k * 32, 1024, 128),
                    sm.read_k_major(sK + k * 32, 1024, SP), k > 0)
    m_row, l_row = {}, {}
    for t in range(nt):                                              # softmax warps (t, q), lane = row
        for q in range(active_warps(S, t)):
            rows = t * 128 + q * 32 + np.arange(32)
            s = tm.ld(q, TM_S + t * TM_S_STRIDE, SP)
            v = s * SCALE_LOG2 + mb[None, :]
            m = v.max(1)
            p = np.exp2((v - m[:, None]).astype(np.float32))
            l = p.sum(1, dtype=np.float32)
            if thresh:
                idx = rows[:, None].astype(np.uint64) * S + np.arange(SP)[None, :]
                p = np.where(DR.keep_mask_np(hkey, None, thresh, idx), p * scale, np.float32(0))
            pb = bf(p)
            for lane in range(32):
                for c in range(0, SP, 8):
                    sm.store_row_units(off_p + t * nchunk * P_CHUNK, q * 32 + lane, c, pb[lane, c:c + 8])
            for lane in range(32):
                m_row[int(rows[lane])], l_row[int(rows[lane])] = m[lane], l[lane]
    for t in range(nt):                                              # MMA warp: O_t = P_t V
        for kk in range(SP // 16):
            tm.umma(TM_O + t * TM_O_STRIDE, HD,
                    sm.read_k_major(off_p + t * nchunk * P_CHUNK + (kk >> 2) * P_CHUNK + (kk & 3) * 32, 1024, 128),
                    sm.read_mn_major(sV + kk * 2048, 8192, 1024, HD), kk > 0)
    ctx, lse = np.full((S, HD), np.nan), np.full(S, np.nan)
    for t in range(nt):                                              # epilogue
        for q in range(active_warps(S, t)):
            o = tm.ld(q, TM_O + t * TM_O_STRIDE, HD)
            for lane in range(32):
                row = t * 128 + q * 32 + lane
                if row < S:
                    ctx[row] = bf(o[lane] * (np.float32(1) / l_row[row]))
                    lse[row] = (m_row[row] + np.log2(l_row[row])) * (1.0 / LOG2E)
    return ctx, lse


def backward_walk(S, SP, Q, K, V, dO, ctx, lse, mask, thresh, scale, hkey, nsplit=2):
    nu, nchunk = (2 if S > 128 else 1), (SP + 63) // 64
    tile, pt = SP * 128, ((SP + 63) // 64) * P_CHUNK
    off_ds, off_pd = 4 * tile, 4 * tile + pt
    sm, tm = Smem(off_pd + pt + 4 * SP * 4 + 128), Tmem()
    TMB_S, TMB_DP, TMB_DV, TMB_DK, TMB_DQ = 0, 192, 0, 64, 384
    sQ, sK, sV, sdO = 0, tile, 2 * tile, 3 * tile
    for base, mat in ((sQ, Q), (sK, K), (sV, V), (sdO, dO)):
        sm.tma_box(base, mat)
    lse2 = np.where(np.arange(SP) < S, np.pad(lse, (0, SP - S)) * LOG2E, np.inf).astype(np.float32)
    delta = np.where(np.arange(SP) < S, np.pad((bf(ctx) * dO[:S]).sum(1), (0, SP - S)), 0.0).astype(np.float32)
    nch = SP // 16
    per = (nch + nsplit - 1) // nsplit                                # 16-column chunks per part
    dqkv = {n: np.full((S, HD), np.nan) for n in ("dq", "dk", "dv")}
    for u in range(nu):
        for k in range(4):                                           # S^T_u = K_u Q^T, dP^T_u = V_u dO^T
            tm.umma(TMB_S, SP, sm.read_k_major(sK + u * 16384 + k * 32, 1024, 128), sm.read_k_major(sQ + k * 32, 1024, SP), k > 0)
        for k in range(4):
            tm.umma(TMB_DP, SP, sm.read_k_major(sV + u * 16384 + k * 32, 1024, 128), sm.read_k_major(sdO + k * 32, 1024, SP), k > 0)
        for q in range(4):                                           # element-wise warps (q, part), lane = key row
            if u * 128 + q * 32 >= S:
                continue
            keys = u * 128 + q * 32 + np.arange(32)
            bias = np.where(keys < S, np.where(np.pad(mask, (0, 256))[keys] != 0, np.float32(0), MASK_LOG2),
                            np.float32(-np.inf)).astype(np.float32)
            for part in range(nsplit):
                c0, c1 = min(part * per, nch) * 16, min((part + 1) * per, nch) * 16
                for c in range(c0, c1, 16):
                    s, dp = tm.ld(q, TMB_S + c, 16), tm.ld(q, TMB_DP + c, 16)
                    p = np.exp2((s * SCALE_LOG2 + bias[:, None]) - lse2[None, c:c + 16]).astype(np.float32)
                    pd = p
                    if thresh:
                        idx = (c + np.arange(16))[None, :].astype(np.uint64) * S + keys[:, None]
                        kp = DR.keep_mask_np(hkey, None, thresh, idx)
                        pd, dp = np.where(kp, p * scale, np.float32(0)), np.where(kp, dp * scale, np.float32(0))
                    ds = p * (dp - delta[None, c:c + 16])            # the 1/8 is applied in the dK / dQ epilogues
                    pdb, dsb = bf(pd), bf(ds)
                    for lane in range(32):
                        for g in range(2):
                            sm.store_row_units(off_pd, q * 32 + lane, c + 8 * g, pdb[lane, 8 * g:8 * g + 8])
                            sm.store_row_units(off_ds, q * 32 + lane, c + 8 * g, dsb[lane, 8 * g:8 * g + 8])
        for kk in range(SP // 16):                                   # dV_u = Pd^T_u dO, dK_u = dS^T_u Q
            a_off = (kk >> 2) * P_CHUNK + (kk & 3) * 32
            tm.umma(TMB_DV, HD, sm.read_k_major(off_pd + a_off, 1024, 128), sm.read_mn_major(sdO + kk * 2048, 8192, 1024, HD), kk > 0)
        for kk in range(SP // 16):
            a_off = (kk >> 2) * P_CHUNK + (kk & 3) * 32
            tm.umma(TMB_DK, HD, sm.read_k_major(off_ds + a_off, 1024, 128), sm.read_mn_major(sQ + kk * 2048, 8192, 1024, HD), kk > 0)
        ksteps = (min(SP, 128) if u == 0 else SP - 128) // 16
        for m in range(nu):                                          # dQ_m += dS[queries m, keys u] K_u
            for kk in range(ksteps):
                tm.umma(TMB_DQ + 64 * m, HD, sm.read_mn_major(off_ds + m * 2 * P_CHUNK + kk * 2048, P_CHUNK, 1024, 128),
                        sm.read_mn_major(sK + u * 16384 + kk * 2048, 8192, 1024, HD), u > 0 or kk > 0)
        for q in range(4):                                           # [dV_u | dK_u] -> dqkv
            if u * 128 + q * 32 >= S:
                continue
            wkv = 128 // nsplit                                      # part -> slice of TMEM columns [dV | dK] = [0,128)
            for part in range(nsplit):
                ckv = part * wkv
                o = tm.ld(q, TMB_DV + ckv, wkv)
                for lane in range(32):
                    key = u * 128 + q * 32 + lane
                    if key < S:
                        dqkv["dv" if ckv < HD else "dk"][key, ckv & (HD - 1):(ckv & (HD - 1)) + wkv] = \
                            bf(o[lane] * np.float32(1.0 if ckv < HD else 0.125))
    for q in range(4):                                               # dQ epilogue: part = 64 / nsplit of the 64 columns
        if q * 32 >= S:
            continue
        wq = HD // nsplit
        for part in range(nsplit):
            o = tm.ld(q, TMB_DQ + part * wq, wq)
            for lane in range(32):
                if q * 32 + lane < S:
                    dqkv["dq"][q * 32 + lane, part * wq:(part + 1) * wq] = bf(o[lane] * np.float32(0.125))
            if nu == 2 and q == 0:
                o = tm.ld(0, TMB_DQ + 64 + part * wq, wq)
                for lane in range(32):
                    if 128 + lane < S:
                        dqkv["dq"][128 + lane, part * wq:(part + 1) * wq] = bf(o[lane] * np.float32(0.125))
    return dqkv["dq"], dqkv["dk"], dqkv["dv"]


@pytest.mark.parametrize("S,p_drop", [(150, 0.0), (150, 0.1), (33, 0.1), (129, 0.0), (16, 0.1), (77, 0.1)])
def test_forward_and_backward_dataflow(S, p_drop):
    SP, Q, K, V, dO, mask, thresh, scale, hkey, keep = make_case(S, 100 + S, p_drop)
    O, lse_ref, dQ, dK, dV = reference(S, Q, K, V, dO, mask, keep, scale)
    ctx, lse = forward_walk(S, SP, Q, K, V, mask, thresh, scale, hkey)
    assert np.isfinite(ctx).all() and np.isfinite(lse).all()
    assert np.abs(ctx - O).max() <= 2e-2                                  # north_star: hidden states 2e-2 abs (bf16)
    assert np.abs(lse - lse_ref).max() <= 2e-3
    for nsplit in (2, 4):
        dq, dk, dv = backward_walk(S, SP, Q, K, V, dO, ctx, lse, mask, thresh, scale, hkey, nsplit)
        for got, ref, name in ((dq, dQ, "dQ"), (dk, dK, "dK"), (dv, dV, "dV")):
            assert np.isfinite(got).all(), (name, nsplit)
            assert np.abs(got - ref).max() <= 1.5e-2 * max(np.abs(dQ).max(), np.abs(dK).max(), np.abs(dV).max()), (name, nsplit)
