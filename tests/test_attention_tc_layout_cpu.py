"""CPU model of the shared-memory operand layouts of csrc/attention_tc.cu (the experimental tcgen05 attention).

The kernels cannot run here (no GPU) and have not run on hardware yet, so this test pins the part that is pure
address arithmetic: it fills a byte-addressed model of shared memory the way TMA (SWIZZLE_128B boxes) and the
element-wise warps (16-byte st.shared at swizzled units) do, then reads every tcgen05.mma operand back through the
canonical UMMA descriptor layouts (K-major / MN-major SWIZZLE_128B as documented in CUTLASS
cute/atom/mma_traits_sm100.hpp: K-major ((8,m),(T,2)):((8T,SBO),(1,T)); MN-major ((T,8,m),(8,k)):((1,T,LBO),(8T,SBO)),
Swizzle<3,4,3> on the byte address) with the start addresses, LBO / SBO and K-step strides the kernels pass, and
checks that each product then computes what the algorithm needs, and that the padding rows an M = 128 operand
drags in (rows past SP, query blocks past the last chunk) stay INSIDE the kernel's shared-memory allocation (their
values are irrelevant: MMA rows are independent and those output rows are never stored).  The constants below
mirror the .cu file; the values are elements tagged by (matrix, row, column) so any mix-up shows."""
import numpy as np
import pytest

P_CHUNK = 128 * 128        # bytes: 128 rows x 64 bf16


def swz(addr):
    """Swizzle<3,4,3>: byte-address bits [4,7) ^= bits [7,10)."""
    return addr ^ (((addr >> 7) & 7) << 4)


class Smem:
    def __init__(self, nbytes):
        self.nbytes = nbytes                                   # the kernel's dynamic allocation past the 1024-byte alignment
        self.v = np.full(nbytes // 2, np.nan, np.float64)      # one slot per bf16 element; NaN = never written

    def _at(self, addr):
        assert 0 <= addr < self.nbytes, f"operand read at byte {addr} outside the {self.nbytes}-byte allocation"
        return self.v[addr // 2]

    def tma_box(self, base, mat):
        """cp.async.bulk.tensor 2-D box [rows][64] with SWIZZLE_128B at a 1024-aligned base."""
        assert base % 1024 == 0 and mat.shape[1] == 64
        for r in range(mat.shape[0]):
            for c in range(64):
                self.v[swz(base + r * 128 + c * 2) // 2] = mat[r, c]

    def store_row_units(self, tile_base, row, col0, vals):
        """The element-wise warps' st.shared.v4: 8 consecutive K elements of `row` starting at col0 (multiple of 8)."""
        assert col0 % 8 == 0 and len(vals) == 8
        addr = tile_base + (col0 >> 6) * P_CHUNK + row * 128 + ((((col0 & 63) >> 3) ^ (row & 7)) << 4)
        for j, x in enumerate(vals):
            self.v[(addr + 2 * j) // 2] = x

    def read_k_major(self, start, sbo, rows, k=16):
        """[rows][16] operand slab of one K = 16 MMA through a K-major SWIZZLE_128B descriptor."""
        out = np.empty((rows, k))
        for mn in range(rows):
            for kk in range(k):
                out[mn, kk] = self._at(swz(start + (mn % 8) * 128 + (mn // 8) * sbo + kk * 2))
        return out

    def read_mn_major(self, start, lbo, sbo, rows, k=16):
        """[rows][16] operand slab (rows = M or N index) through an MN-major SWIZZLE_128B descriptor."""
        out = np.empty((rows, k))
        for mn in range(rows):
            for kk in range(k):
                out[mn, kk] = self._at(swz(start + (mn % 64) * 2 + (mn // 64) * lbo + (kk % 8) * 128 + (kk // 8) * sbo))
        return out


def tagged(tag, rows, cols=64):
    r, c = np.meshgrid(np.arange(rows), np.arange(cols), indexing="ij")
    return tag * 1e6 + r * 1e3 + c


@pytest.mark.parametrize("S", [160, 159, 150, 145, 144, 129, 128, 127, 100, 97, 96, 76, 65, 64, 33, 17, 16, 15, 1])
def test_forward_operands(S):
    SP = (S + 15) // 16 * 16
    nt = 2 if S > 128 else 1
    nchunk = (SP + 63) // 64
    tile = SP * 128
    off_p = 2 * 3 * tile
    sm = Smem(off_p + nt * nchunk * P_CHUNK + 2 * SP * 4 + 128)        # tc_smem(): ... + mask rows + barriers
    Q, K, V = tagged(1, SP), tagged(2, SP), tagged(3, SP)
    for buf in range(2):
        base = buf * 3 * tile
        sm.tma_box(base, Q); sm.tma_box(base + tile, K); sm.tma_box(base + 2 * tile, V)
        sQ, sK, sV = base, base + tile, base + 2 * tile
        # S_t = Q_t K^T: 4 K steps of 16 over the head dimension
        for t in range(nt):
            rows = min(128, SP - 128 * t)
            for k in range(4):
                a = sm.read_k_major(sQ + t * 16384 + k * 32, 1024, 128)
                b = sm.read_k_major(sK + k * 32, 1024, SP)
                assert np.array_equal(a[:rows], Q[128 * t:128 * t + rows, 16 * k:16 * k + 16])
                assert np.array_equal(b, K[:, 16 * k:16 * k + 16])
    # P_t written by the softmax threads, read back as the K-major A operand of O_t = P_t V (K = keys)
    P = [tagged(4 + t, 128, SP) for t in range(nt)]
    for t in range(nt):
        sP = off_p + t * nchunk * P_CHUNK
        for row in range(128):
            for c in range(0, SP, 8):
                sm.store_row_units(sP, row, c, P[t][row, c:c + 8])
        for kk in range(SP // 16):
            a = sm.read_k_major(sP + (kk >> 2) * P_CHUNK + (kk & 3) * 32, 1024, 128)
            assert np.array_equal(a, P[t][:, 16 * kk:16 * kk + 16])
            # V as the MN-major B operand: N = head dim (64), K = keys 16 kk .. 16 kk + 15
            b = sm.read_mn_major(sV + kk * 2048, 8192, 1024, 64)
            assert np.array_equal(b, V[16 * kk:16 * kk + 16, :].T)


@pytest.mark.parametrize("S", [160, 159, 150, 145, 144, 129, 128, 127, 100, 97, 96, 76, 65, 64, 33, 17, 16, 15, 1])
def test_backward_operands(S):
    SP = (S + 15) // 16 * 16
    nu = 2 if S > 128 else 1
    nchunk = (SP + 63) // 64
    tile = SP * 128
    pt = nchunk * P_CHUNK
    off_ds, off_pd = 4 * tile, 4 * tile + pt
    total = off_pd + pt + 4 * SP * 4 + 128
    sm = Smem(total)
    Q, K, V, dO = tagged(1, SP), tagged(2, SP), tagged(3, SP), tagged(4, SP)
    sQ, sK, sV, sdO = 0, tile, 2 * tile, 3 * tile
    for base, m in ((sQ, Q), (sK, K), (sV, V), (sdO, dO)):
        sm.tma_box(base, m)
    for u in range(nu):
        keys = min(128, SP - 128 * u)
        # S^T_u = K_u Q^T and dP^T_u = V_u dO^T
        for k in range(4):
            a = sm.read_k_major(sK + u * 16384 + k * 32, 1024, 128)
            assert np.array_equal(a[:keys], K[128 * u:128 * u + keys, 16 * k:16 * k + 16])
            assert np.array_equal(sm.read_k_major(sQ + k * 32, 1024, SP), Q[:, 16 * k:16 * k + 16])
            a = sm.read_k_major(sV + u * 16384 + k * 32, 1024, 128)
            assert np.array_equal(a[:keys], V[128 * u:128 * u + keys, 16 * k:16 * k + 16])
            assert np.array_equal(sm.read_k_major(sdO + k * 32, 1024, SP), dO[:, 16 * k:16 * k + 16])
        # the element-wise warps write Pd^T_u and dS^T_u: row = key inside the tile, columns = queries
        PdT, dST = tagged(5 + u, 128, SP), tagged(7 + u, 128, SP)
        for row in range(keys if keys % 32 == 0 else (keys + 31) // 32 * 32):   # whole active warps write
            for c in range(0, SP, 8):
                sm.store_row_units(off_pd, row, c, PdT[row, c:c + 8])
                sm.store_row_units(off_ds, row, c, dST[row, c:c + 8])
        for kk in range(SP // 16):       # dV_u = Pd^T_u dO, dK_u = dS^T_u Q: K = queries
            a_off = (kk >> 2) * P_CHUNK + (kk & 3) * 32
            assert np.array_equal(sm.read_k_major(off_pd + a_off, 1024, 128)[:keys], PdT[:keys, 16 * kk:16 * kk + 16])
            assert np.array_equal(sm.read_k_major(off_ds + a_off, 1024, 128)[:keys], dST[:keys, 16 * kk:16 * kk + 16])
            assert np.array_equal(sm.read_mn_major(sdO + kk * 2048, 8192, 1024, 64), dO[16 * kk:16 * kk + 16].T)
            assert np.array_equal(sm.read_mn_major(sQ + kk * 2048, 8192, 1024, 64), Q[16 * kk:16 * kk + 16].T)
        # dQ_m += dS[queries of tile m, keys of tile u] K_u: A = dS^T tile MN-major, B = K_u MN-major, K = keys
        for m in range(nu):
            qrows = min(128, SP - 128 * m)
            for kk in range(keys // 16):
                a = sm.read_mn_major(off_ds + m * 2 * P_CHUNK + kk * 2048, P_CHUNK, 1024, 128)
                assert np.array_equal(a[:qrows], dST[16 * kk:16 * kk + 16, 128 * m:128 * m + qrows].T)
                b = sm.read_mn_major(sK + u * 16384 + kk * 2048, 8192, 1024, 64)
                assert np.array_equal(b, K[128 * u + 16 * kk:128 * u + 16 * kk + 16].T)


def test_switch_is_off_by_default_and_toggles(monkeypatch):
    """The experimental kernels must never be the default path: the switch starts off (no UC2_ATTN_TCGEN05 in the
    environment of this test run) and uc2_attention_tc_enable returns the previous setting."""
    import os
    from uc2_b200 import _lib
    if os.environ.get("UC2_ATTN_TCGEN05") == "1":
        pytest.skip("switch forced on through the environment")
    assert _lib.attention_tc_enabled() is False
    assert _lib.lib().uc2_attention_tc_enable(1) == 0
    assert _lib.attention_tc_enabled() is True
    assert _lib.lib().uc2_attention_tc_enable(0) == 1
    assert _lib.attention_tc_enabled() is False
