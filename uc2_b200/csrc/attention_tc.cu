// tcgen05 / TMEM forward attention for S <= 160 (every BASELINE training shape but VTLM), 12 heads x 64.
// Replaces the same reference code as attention.cu (BertSelfAttention.forward model/layer.py:80-100) with the
// two products on the 5th-generation tensor cores instead of mma.sync:
//
//   persistent CTAs (one per SM) loop over (batch, head) items; per item
//     warp 0      TMA producer: Q, K, V of the head as three [SP x 64] SWIZZLE_128B boxes of the packed qkv
//                 tensor (double buffered: the next item's tiles land under this item's softmax), and the
//                 additive key-mask row in the exp2 domain
//     warp 1      one thread issues  S_t = Q_t K^T  (M 128, N SP, K 64; t = query tile 0 / 1) into TMEM, and after
//                 the softmax warps published P_t,  O_t = P_t V  (M 128, N 64, K SP; V is the MN-major B operand
//                 straight from its [key][d] rows)
//     warps 2-9   two softmax groups, group t owns query tile t: a thread owns ONE query row (TMEM lane), so row
//                 max / row sum need no shuffles; two passes over the row in TMEM (max, then exp2 + sum + the
//                 counter-hash dropout of attention.cu), P written as bf16 into the K-major SWIZZLE_128B
//                 shared-memory layout the MMA's A descriptor reads; then the epilogue O / rowsum -> ctx, lse.
//   TMEM: S_0 [0,160) S_1 [192,352) O_0 [384,448) O_1 [448,512) of 512 columns.
//
// STATUS: written and compiled (ptxas / SASS checked) at the end of round 1 after the round's GPU budget was
// spent -- NOT YET RUN ON HARDWARE.  It is therefore off by default: uc2_attention_fwd(_dropout) only route here
// after uc2_attention_tc_enable(1) or with UC2_ATTN_TCGEN05=1 in the environment, and its parity test
// (tests/test_attention_tc_gpu.py) runs only with UC2_TEST_EXPERIMENTAL=1.  Results are defined to be those of
// attention_fwd_bh_kernel (same masks, same dropout stream, same lse), so the existing backward pairs with it.
// Checked on the CPU meanwhile (tests/test_attention_tc_{layout,protocol,dataflow}_cpu.py): every MMA operand as
// read through its descriptor, the mbarrier protocol under random interleavings, and the data path on real numbers.
// If you change an offset, a stride, a barrier count or a parity here, change the mirrored constant there.
#include <atomic>
#include <mutex>
#include <stdlib.h>

#include "common.cuh"
#include "ptx.cuh"

namespace uc2 {
namespace {

constexpr int HD = 64;
constexpr int NH = 12;
constexpr int QKV_LD = 3 * HID;
constexpr int TC_MAX_SP = 160;
constexpr int TC_THREADS = 320;              // TMA warp, MMA warp, 2 x 4 softmax warps
constexpr float LOG2E = 1.4426950408889634f;
constexpr float SCALE_LOG2 = 0.125f * LOG2E;
constexpr float MASK_LOG2 = -10000.0f * LOG2E;
// accumulator column bases kept on multiples of 64 (the 160-column S tiles get 192-column slots)
constexpr uint32_t TM_S = 0, TM_S_STRIDE = 192, TM_O = 384, TM_O_STRIDE = 64, TMEM_COLS = 512;
constexpr uint32_t P_CHUNK_BYTES = 128 * 128;   // 128 query rows x 64 keys of bf16

struct TcParams {
    const long long* mask;
    bf16* ctx;
    float* lse;
    int B, S, SP, items;
    DropCfg drop;
};

// shared-memory plan shared by host and device
struct TcSmem {
    uint32_t tile_bytes, buf_bytes, p_tile_bytes, off_p, off_mbias, off_bar, total;
    int nt, nchunk;
};
__host__ __device__ inline TcSmem tc_smem(int S, int SP) {
    TcSmem L;
    L.nt = S > 128 ? 2 : 1;
    L.nchunk = (SP + 63) >> 6;
    L.tile_bytes = SP * 128u;
    L.buf_bytes = 3u * L.tile_bytes;
    L.p_tile_bytes = L.nchunk * P_CHUNK_BYTES;
    L.off_p = 2u * L.buf_bytes;
    L.off_mbias = L.off_p + L.nt * L.p_tile_bytes;
    L.off_bar = L.off_mbias + 2u * SP * 4u;
    L.total = 1024u + L.off_bar + 128u;
    return L;
}

__host__ __device__ inline int active_warps(int S, int t) {
    const int rows = S - 128 * t;
    return rows <= 0 ? 0 : (rows >= 128 ? 4 : (rows + 31) >> 5);
}

// One pass-2 chunk: NC (16 or 32) score columns of this thread's row -> probabilities -> bf16 into the P tile
template <int NC>
__device__ __forceinline__ void softmax_chunk(const uint32_t* r, const float* mb, int c, float m, float& l,
                                              const DropCfg& drop, uint32_t hkey, uint32_t row_idx0, uint32_t prow,
                                              uint32_t sw) {
    uint32_t pk[NC / 2];
#pragma unroll
    for (int j = 0; j < NC; j += 4) {
        const float4 bb = *reinterpret_cast<const float4*>(mb + c + j);
        float p0 = fast_ex2(fmaf(__uint_as_float(r[j]), SCALE_LOG2, bb.x) - m);
        float p1 = fast_ex2(fmaf(__uint_as_float(r[j + 1]), SCALE_LOG2, bb.y) - m);
        float p2 = fast_ex2(fmaf(__uint_as_float(r[j + 2]), SCALE_LOG2, bb.z) - m);
        float p3 = fast_ex2(fmaf(__uint_as_float(r[j + 3]), SCALE_LOG2, bb.w) - m);
        l += (p0 + p1) + (p2 + p3);
        if (drop.thresh) {
            // layer.py:94: the normaliser keeps every key, the dropped and rescaled probabilities only enter P.V
            bool k0, k1, k2, k3;
            drop_keep2(hkey, row_idx0 + c + j, drop.thresh, k0, k1);
            drop_keep2(hkey, row_idx0 + c + j + 2, drop.thresh, k2, k3);
            p0 = k0 ? p0 * drop.scale : 0.f; p1 = k1 ? p1 * drop.scale : 0.f;
            p2 = k2 ? p2 * drop.scale : 0.f; p3 = k3 ? p3 * drop.scale : 0.f;
        }
        pk[j / 2] = pack_bf16(p0, p1);
        pk[j / 2 + 1] = pack_bf16(p2, p3);
    }
#pragma unroll
    for (int g = 0; g < NC / 8; ++g) {
        const uint32_t key0 = c + 8 * g;                       // 8 keys = one 16-byte unit of the swizzled row
        const uint32_t addr = prow + (key0 >> 6) * P_CHUNK_BYTES + ((((key0 & 63u) >> 3) ^ sw) << 4);
        ptx::st_shared_v4(addr, pk[4 * g], pk[4 * g + 1], pk[4 * g + 2], pk[4 * g + 3]);
    }
}

template <int NC>
__device__ __forceinline__ float max_chunk(const uint32_t* r, const float* mb, int c, float m) {
#pragma unroll
    for (int j = 0; j < NC; j += 4) {
        const float4 bb = *reinterpret_cast<const float4*>(mb + c + j);
        m = fmaxf(m, fmaxf(fmaxf(fmaf(__uint_as_float(r[j]), SCALE_LOG2, bb.x),
                                 fmaf(__uint_as_float(r[j + 1]), SCALE_LOG2, bb.y)),
                           fmaxf(fmaf(__uint_as_float(r[j + 2]), SCALE_LOG2, bb.z),
                                 fmaf(__uint_as_float(r[j + 3]), SCALE_LOG2, bb.w))));
    }
    return m;
}

__global__ void __launch_bounds__(TC_THREADS, 1)
attention_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmap_qkv, const TcParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* gen = smem_raw + (base - ptx::smem_u32(smem_raw));
    const int S = p.S, SP = p.SP;
    const TcSmem L = tc_smem(S, SP);
    float* mbias = reinterpret_cast<float*>(gen + L.off_mbias);
    const uint32_t bar = base + L.off_bar;
    // barriers (8 B each): full[2] empty[2] s_full[2] p_full[2] o_full[2] o_empty[2], then the TMEM base address
    auto full_bar = [&](int b) { return bar + 8u * b; };
    auto empty_bar = [&](int b) { return bar + 16u + 8u * b; };
    auto sfull_bar = [&](int t) { return bar + 32u + 8u * t; };
    auto pfull_bar = [&](int t) { return bar + 48u + 8u * t; };
    auto ofull_bar = [&](int t) { return bar + 64u + 8u * t; };
    auto oempty_bar = [&](int t) { return bar + 80u + 8u * t; };
    const uint32_t tmem_ptr_addr = bar + 96u;
    volatile uint32_t* tmem_ptr_gen = reinterpret_cast<volatile uint32_t*>(gen + L.off_bar + 96);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        ptx::prefetch_tmap(&tmap_qkv);
        for (int b = 0; b < 2; ++b) {
            ptx::mbar_init(full_bar(b), 2);            // TMA transaction arrive + mask-row arrive
            ptx::mbar_init(empty_bar(b), 1);
        }
        for (int t = 0; t < 2; ++t) {
            const int na = active_warps(S, t);
            ptx::mbar_init(sfull_bar(t), 1);
            ptx::mbar_init(pfull_bar(t), na > 0 ? na : 1);
            ptx::mbar_init(ofull_bar(t), 1);
            ptx::mbar_init(oempty_bar(t), na > 0 ? na : 1);
        }
        ptx::fence_barrier_init();
        ptx::fence_proxy_async();
    }
    if (warp == 1) {
        ptx::tmem_alloc(tmem_ptr_addr, TMEM_COLS);
        ptx::tmem_relinquish();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_gen;
    griddep_sync();        // everything above may run under the tail of the QKV GEMM

    if (warp == 0) {
        // ===================================== producer =====================================
        int it = 0;
        for (int item = blockIdx.x; item < p.items; item += gridDim.x, ++it) {
            const int buf = it & 1;
            const uint32_t ph = (it >> 1) & 1u;
            const int b = item / NH, h = item - b * NH;
            if (lane == 0) ptx::mbar_wait(empty_bar(buf), ph ^ 1u);   // the MMAs of item it - 2 retired
            __syncwarp();
            if (lane == 0) {
                const uint32_t dst = base + buf * L.buf_bytes;
                ptx::mbar_arrive_expect_tx(full_bar(buf), L.buf_bytes);
                ptx::tma_load_2d(dst, &tmap_qkv, full_bar(buf), h * HD, b * S);
                ptx::tma_load_2d(dst + L.tile_bytes, &tmap_qkv, full_bar(buf), HID + h * HD, b * S);
                ptx::tma_load_2d(dst + 2u * L.tile_bytes, &tmap_qkv, full_bar(buf), 2 * HID + h * HD, b * S);
            }
            // additive key mask of model.py:433-436 in the exp2 domain; padding columns [S, SP) never contribute
            float* mb = mbias + buf * SP;
            const long long* mrow = p.mask + (long long)b * S;
            for (int j = lane; j < SP; j += 32) mb[j] = j < S ? (mrow[j] != 0 ? 0.f : MASK_LOG2) : -INFINITY;
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(full_bar(buf));
        }
    } else if (warp == 1) {
        // ===================================== MMA issuer ===================================
        if (lane == 0) {
            const uint32_t idesc_s = ptx::idesc_bf16_f32(128, SP, false, false);
            const uint32_t idesc_o = ptx::idesc_bf16_f32(128, HD, false, true);
            int it = 0;
            for (int item = blockIdx.x; item < p.items; item += gridDim.x, ++it) {
                const int buf = it & 1;
                const uint32_t ph = (it >> 1) & 1u, itp = it & 1u;
                ptx::mbar_wait(full_bar(buf), ph);
                ptx::tc_fence_after();
                const uint32_t sQ = base + buf * L.buf_bytes, sK = sQ + L.tile_bytes, sV = sK + L.tile_bytes;
                // S_t = Q_t K^T.  The S columns are free: p_full(t) of the previous item was waited on below.
                // Tile 1 reads 128 rows from Q row 128 on; rows past SP are the K tile (finite, never stored).
                for (int t = 0; t < L.nt; ++t) {
#pragma unroll
                    for (int k = 0; k < HD / 16; ++k) {
                        const uint64_t da = ptx::smem_desc_sw128(sQ + t * 16384u + k * 32u, 16u, 1024u);
                        const uint64_t db = ptx::smem_desc_sw128(sK + k * 32u, 16u, 1024u);
                        ptx::umma_bf16(tmem_base + TM_S + t * TM_S_STRIDE, da, db, idesc_s, k > 0 ? 1u : 0u);
                    }
                    ptx::umma_commit(sfull_bar(t));
                }
                // O_t = P_t V once group t has written P_t and read the previous item's O_t
                for (int t = 0; t < L.nt; ++t) {
                    ptx::mbar_wait(pfull_bar(t), itp);
                    ptx::mbar_wait(oempty_bar(t), itp ^ 1u);
                    ptx::tc_fence_after();
                    const uint32_t sP = base + L.off_p + t * L.p_tile_bytes;
                    for (int kk = 0; kk < SP / 16; ++kk) {
                        const uint64_t da =
                            ptx::smem_desc_sw128(sP + (kk >> 2) * P_CHUNK_BYTES + (kk & 3) * 32u, 16u, 1024u);
                        const uint64_t db = ptx::smem_desc_sw128(sV + kk * 2048u, 8192u, 1024u);
                        ptx::umma_bf16(tmem_base + TM_O + t * TM_O_STRIDE, da, db, idesc_o, kk > 0 ? 1u : 0u);
                    }
                    ptx::umma_commit(ofull_bar(t));
                }
                ptx::umma_commit(empty_bar(buf));       // Q, K, V of this buffer are dead once all of the above retire
            }
        }
    } else {
        // ===================================== softmax groups ================================
        const int t = (warp - 2) >> 2;              // query tile of this group
        const int q = warp & 3;                     // TMEM lane quarter this warp may access
        if (t * 128 + q * 32 < S) {
            const int row = t * 128 + q * 32 + lane;                  // query row inside the (batch, head)
            const bool row_ok = row < S;
            const uint32_t lane_sel = static_cast<uint32_t>(q * 32) << 16;
            const uint32_t tS = tmem_base + lane_sel + TM_S + t * TM_S_STRIDE;
            const uint32_t tO = tmem_base + lane_sel + TM_O + t * TM_O_STRIDE;
            const uint32_t prow = base + L.off_p + t * L.p_tile_bytes + static_cast<uint32_t>(q * 32 + lane) * 128u;
            const uint32_t sw = static_cast<uint32_t>(lane & 7);
            const uint32_t row_idx0 = static_cast<uint32_t>(row) * static_cast<uint32_t>(S);
            int it = 0;
            for (int item = blockIdx.x; item < p.items; item += gridDim.x, ++it) {
                const int buf = it & 1;
                const uint32_t ph = (it >> 1) & 1u, itp = it & 1u;
                const int b = item / NH, h = item - b * NH;
                const float* mb = mbias + buf * SP;
                const uint32_t hkey = drop_head_key(p.drop.key, b * NH + h);
                ptx::mbar_wait(full_bar(buf), ph);                    // the mask row is there
                ptx::mbar_wait(sfull_bar(t), itp);                    // ... and so is S_t (and P_t is free again)
                ptx::tc_fence_after();
                uint32_t r[32];
                float m = -INFINITY;
                int c = 0;
                for (; c + 32 <= SP; c += 32) {
                    ptx::tmem_ld_32x32(tS + c, r);
                    ptx::tmem_wait_ld();
                    m = max_chunk<32>(r, mb, c, m);
                }
                if (c < SP) {
                    ptx::tmem_ld_32x16(tS + c, r);
                    ptx::tmem_wait_ld();
                    m = max_chunk<16>(r, mb, c, m);
                }
                float l = 0.f;
                for (c = 0; c + 32 <= SP; c += 32) {
                    ptx::tmem_ld_32x32(tS + c, r);
                    ptx::tmem_wait_ld();
                    softmax_chunk<32>(r, mb, c, m, l, p.drop, hkey, row_idx0, prow, sw);
                }
                if (c < SP) {
                    ptx::tmem_ld_32x16(tS + c, r);
                    ptx::tmem_wait_ld();
                    softmax_chunk<16>(r, mb, c, m, l, p.drop, hkey, row_idx0, prow, sw);
                }
                ptx::fence_proxy_async();        // P_t: generic-proxy stores -> visible to the tensor core's reads
                ptx::tc_fence_before();          // S_t reads are complete (wait::ld above) before the barrier
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive(pfull_bar(t));

                // epilogue: O_t / rowsum -> ctx (heads merged: layer.py:98-100 is free), lse for the backward
                ptx::mbar_wait(ofull_bar(t), itp);
                ptx::tc_fence_after();
                const float inv = 1.f / l;
                bf16* orow = p.ctx + ((long long)b * S + row) * HID + h * HD;
#pragma unroll
                for (int c2 = 0; c2 < HD; c2 += 32) {
                    ptx::tmem_ld_32x32(tO + c2, r);
                    ptx::tmem_wait_ld();
                    if (row_ok) {
#pragma unroll
                        for (int g = 0; g < 2; ++g) {
                            uint32_t o[8];
#pragma unroll
                            for (int j = 0; j < 8; ++j)
                                o[j] = pack_bf16(__uint_as_float(r[16 * g + 2 * j]) * inv,
                                                 __uint_as_float(r[16 * g + 2 * j + 1]) * inv);
                            ptx::stg256(orow + c2 + 16 * g, o[0], o[1], o[2], o[3], o[4], o[5], o[6], o[7]);
                        }
                    }
                }
                if (row_ok) p.lse[((long long)b * NH + h) * S + row] = (m + log2f(l)) * (1.f / LOG2E);
                ptx::tc_fence_before();
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive(oempty_bar(t));
            }
        }
    }

    // ------------------------------------------- teardown -------------------------------------------
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

// =================================================================================================
// Backward, same structure, KEY-major: the recomputed scores live in TMEM as S^T (lane = key, column = query), so
// the element-wise stage needs no row reductions (lse and delta = rowsum(dO * O) are per-COLUMN vectors in shared
// memory, the key's mask bias is a per-thread scalar) and both halves of the query columns of a key row can go to
// two different warps.  Per (batch, head) item and key tile u (128 keys; u = 1 holds keys 128..159):
//     S^T_u = K_u Q^T, dP^T_u = V_u dO^T                                   (M 128, N SP, K 64; TMEM [0,160) [192,352))
//     element-wise: P = exp2(s - lse), dropout, dS = P (dP - delta)       -> Pd^T_u, dS^T_u as bf16 K-major tiles
//     dV_u = Pd^T_u dO, dK_u = dS^T_u Q / 8                                (M 128, N 64, K SP; alias TMEM [0,64) [64,128))
//     dQ_m += dS_u K_u / 8  for the query tiles m                         (A = the dS^T tile read MN-major; TMEM [384,512))
// Tiles of Q, K, V, dO come by TMA once per item (single buffered: the next item's loads fly under this item's
// last epilogues); 8 element-wise warps = 4 TMEM lane quarters x 2 column halves.
// =================================================================================================
struct TcBwdParams {
    const long long* mask;
    const bf16* ctx;
    const bf16* dctx;
    const float* lse;
    bf16* dqkv;
    int B, S, SP, items;
    DropCfg drop;
};

struct TcBwdSmem {
    uint32_t tile_bytes, pt_bytes, off_ds, off_pd, off_vec, off_bar, total;
    int nu, nchunk;
};
__host__ __device__ inline TcBwdSmem tc_bwd_smem(int S, int SP) {
    TcBwdSmem L;
    L.nu = S > 128 ? 2 : 1;
    L.nchunk = (SP + 63) >> 6;
    L.tile_bytes = SP * 128u;
    L.pt_bytes = L.nchunk * P_CHUNK_BYTES;
    L.off_ds = 4u * L.tile_bytes;            // [Q][K][V][dO] then dS^T, then Pd^T (so MN-major over-reads of dS^T
    L.off_pd = L.off_ds + L.pt_bytes;        // for the padding query blocks of dQ stay inside the allocation)
    L.off_vec = L.off_pd + L.pt_bytes;       // 2 buffers x (lse2[SP], delta[SP])
    L.off_bar = L.off_vec + 4u * SP * 4u;
    L.total = 1024u + L.off_bar + 128u;
    return L;
}

constexpr uint32_t TMB_S = 0, TMB_DP = 192, TMB_DV = 0, TMB_DK = 64, TMB_DQ = 384;

// W (16, 32 or 64) accumulator columns of this thread's row, times `scale`, -> W bf16 at dst
template <int W>
__device__ __forceinline__ void store_cols(uint32_t taddr, bf16* dst, bool ok, float scale) {
    uint32_t r[16];
#pragma unroll
    for (int c = 0; c < W; c += 16) {
        ptx::tmem_ld_32x16(taddr + c, r);
        ptx::tmem_wait_ld();
        if (ok) {
            uint32_t o[8];
#pragma unroll
            for (int j = 0; j < 8; ++j)
                o[j] = pack_bf16(__uint_as_float(r[2 * j]) * scale, __uint_as_float(r[2 * j + 1]) * scale);
            ptx::stg256(dst + c, o[0], o[1], o[2], o[3], o[4], o[5], o[6], o[7]);
        }
    }
}

// NSPLIT element-wise warps share a TMEM lane quarter (2: 8 warps, 4: 16 warps = four per scheduler; the loops
// are bound by instruction issue and latency, like the GELU epilogues of the GEMM that run 16 warps for this reason)
__host__ __device__ constexpr int bwd_threads(int nsplit) { return 64 + 128 * nsplit; }

template <int NSPLIT>
__global__ void __launch_bounds__(bwd_threads(NSPLIT), 1)
attention_bwd_tc_kernel(const __grid_constant__ CUtensorMap tmap_qkv, const __grid_constant__ CUtensorMap tmap_do,
                        const TcBwdParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* gen = smem_raw + (base - ptx::smem_u32(smem_raw));
    const int S = p.S, SP = p.SP;
    const TcBwdSmem L = tc_bwd_smem(S, SP);
    float* vec = reinterpret_cast<float*>(gen + L.off_vec);          // [buf][lse2 SP | delta SP]
    const uint32_t bar = base + L.off_bar;
    // barriers (8 B each): ld_full ld_empty sd_full[2] pds_full[2] kv_full[2] s_empty[2] dq_full dq_empty, TMEM ptr
    const uint32_t ld_full = bar, ld_empty = bar + 8u, dq_full = bar + 80u, dq_empty = bar + 88u;
    auto sd_full = [&](int u) { return bar + 16u + 8u * u; };
    auto pds_full = [&](int u) { return bar + 32u + 8u * u; };
    auto kv_full = [&](int u) { return bar + 48u + 8u * u; };
    auto s_empty = [&](int u) { return bar + 64u + 8u * u; };
    const uint32_t tmem_ptr_addr = bar + 96u;
    volatile uint32_t* tmem_ptr_gen = reinterpret_cast<volatile uint32_t*>(gen + L.off_bar + 96);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nu = L.nu;

    if (threadIdx.x == 0) {
        ptx::prefetch_tmap(&tmap_qkv);
        ptx::prefetch_tmap(&tmap_do);
        ptx::mbar_init(ld_full, 1);
        ptx::mbar_init(ld_empty, 1);
        for (int u = 0; u < 2; ++u) {
            const int na = NSPLIT * active_warps(S, u);       // NSPLIT column parts per active lane quarter
            ptx::mbar_init(sd_full(u), 1);
            ptx::mbar_init(pds_full(u), na > 0 ? na : 1);
            ptx::mbar_init(kv_full(u), 1);
            ptx::mbar_init(s_empty(u), na > 0 ? na : 1);
        }
        ptx::mbar_init(dq_full, 1);
        ptx::mbar_init(dq_empty, NSPLIT * active_warps(S, 0));
        ptx::fence_barrier_init();
        ptx::fence_proxy_async();
    }
    if (warp == 1) {
        ptx::tmem_alloc(tmem_ptr_addr, TMEM_COLS);
        ptx::tmem_relinquish();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_gen;
    griddep_sync();

    const uint32_t sQ = base, sK = sQ + L.tile_bytes, sV = sK + L.tile_bytes, sdO = sV + L.tile_bytes;
    const uint32_t sDS = base + L.off_ds, sPD = base + L.off_pd;

    if (warp == 0) {
        // ===================================== producer =====================================
        if (lane == 0) {
            int it = 0;
            for (int item = blockIdx.x; item < p.items; item += gridDim.x, ++it) {
                const int b = item / NH, h = item - b * NH;
                ptx::mbar_wait(ld_empty, (it & 1u) ^ 1u);          // every MMA of the previous item retired
                ptx::mbar_arrive_expect_tx(ld_full, 4u * L.tile_bytes);
                ptx::tma_load_2d(sQ, &tmap_qkv, ld_full, h * HD, b * S);
                ptx::tma_load_2d(sK, &tmap_qkv, ld_full, HID + h * HD, b * S);
                ptx::tma_load_2d(sV, &tmap_qkv, ld_full, 2 * HID + h * HD, b * S);
                ptx::tma_load_2d(sdO, &tmap_do, ld_full, h * HD, b * S);
            }
        }
    } else if (warp == 1) {
        // ===================================== MMA issuer ===================================
        if (lane == 0) {
            const uint32_t idesc_s = ptx::idesc_bf16_f32(128, SP, false, false);
            const uint32_t idesc_kv = ptx::idesc_bf16_f32(128, HD, false, true);
            const uint32_t idesc_dq = ptx::idesc_bf16_f32(128, HD, true, true);
            int it = 0;
            for (int item = blockIdx.x; item < p.items; item += gridDim.x, ++it) {
                const uint32_t itp = it & 1u;
                ptx::mbar_wait(ld_full, itp);
                ptx::tc_fence_after();
                for (int u = 0; u < nu; ++u) {
                    // TMEM [0,352) is free once the dV / dK of the previous key tile have been read out
                    if (u == 0) ptx::mbar_wait(s_empty(nu - 1), itp ^ 1u);
                    else ptx::mbar_wait(s_empty(u - 1), itp);
                    ptx::tc_fence_after();
                    // key tile 1 reads 128 rows from row 128 of K / V on; rows past SP are the following tiles
#pragma unroll
                    for (int k = 0; k < HD / 16; ++k)
                        ptx::umma_bf16(tmem_base + TMB_S, ptx::smem_desc_sw128(sK + u * 16384u + k * 32u, 16u, 1024u),
                                       ptx::smem_desc_sw128(sQ + k * 32u, 16u, 1024u), idesc_s, k > 0 ? 1u : 0u);
#pragma unroll
                    for (int k = 0; k < HD / 16; ++k)
                        ptx::umma_bf16(tmem_base + TMB_DP, ptx::smem_desc_sw128(sV + u * 16384u + k * 32u, 16u, 1024u),
                                       ptx::smem_desc_sw128(sdO + k * 32u, 16u, 1024u), idesc_s, k > 0 ? 1u : 0u);
                    ptx::umma_commit(sd_full(u));

                    ptx::mbar_wait(pds_full(u), itp);              // Pd^T_u, dS^T_u written; S^T, dP^T read
                    if (u == 0) ptx::mbar_wait(dq_empty, itp ^ 1u);   // previous item's dQ read out
                    ptx::tc_fence_after();
                    for (int kk = 0; kk < SP / 16; ++kk) {         // K dimension = queries
                        const uint32_t a_off = (kk >> 2) * P_CHUNK_BYTES + (kk & 3) * 32u;
                        ptx::umma_bf16(tmem_base + TMB_DV, ptx::smem_desc_sw128(sPD + a_off, 16u, 1024u),
                                       ptx::smem_desc_sw128(sdO + kk * 2048u, 8192u, 1024u), idesc_kv, kk > 0 ? 1u : 0u);
                    }
                    for (int kk = 0; kk < SP / 16; ++kk) {
                        const uint32_t a_off = (kk >> 2) * P_CHUNK_BYTES + (kk & 3) * 32u;
                        ptx::umma_bf16(tmem_base + TMB_DK, ptx::smem_desc_sw128(sDS + a_off, 16u, 1024u),
                                       ptx::smem_desc_sw128(sQ + kk * 2048u, 8192u, 1024u), idesc_kv, kk > 0 ? 1u : 0u);
                    }
                    // dQ_m += dS[queries of tile m, keys of tile u] K_u: A is the dS^T tile read MN-major (64-query
                    // blocks P_CHUNK_BYTES apart, 16 key rows per K step), B = K rows of tile u, MN-major
                    const int ksteps = (u == 0 ? (SP < 128 ? SP : 128) : SP - 128) / 16;
                    for (int m = 0; m < nu; ++m)
                        for (int kk = 0; kk < ksteps; ++kk)
                            ptx::umma_bf16(tmem_base + TMB_DQ + m * 64u,
                                           ptx::smem_desc_sw128(sDS + m * 2u * P_CHUNK_BYTES + kk * 2048u, P_CHUNK_BYTES, 1024u),
                                           ptx::smem_desc_sw128(sK + u * 16384u + kk * 2048u, 8192u, 1024u), idesc_dq,
                                           (u > 0 || kk > 0) ? 1u : 0u);
                    ptx::umma_commit(kv_full(u));
                }
                ptx::umma_commit(ld_empty);
                ptx::umma_commit(dq_full);
            }
        }
    } else {
        // ===================================== element-wise warps ============================
        const int q = warp & 3;                         // TMEM lane quarter
        const int part = (warp - 2) >> 2;               // which part of the query columns / of the outputs
        const int ew = warp - 2;
        const uint32_t lane_sel = static_cast<uint32_t>(q * 32) << 16;
        const int kr = q * 32 + lane;                   // row inside a 128-row tile
        const uint32_t sw = static_cast<uint32_t>(lane & 7);
        const int nch = SP / 16, per = (nch + NSPLIT - 1) / NSPLIT;          // 16-column chunks per part
        const int c_begin = min(part * per, nch) * 16, c_end = min((part + 1) * per, nch) * 16;
        int it = 0;
        for (int item = blockIdx.x; item < p.items; item += gridDim.x, ++it) {
            const uint32_t itp = it & 1u;
            const int b = item / NH, h = item - b * NH;
            const long long row0 = (long long)b * S;
            // ---- per-query vectors: lse in the exp2 domain (+inf for padding columns -> P = 0), delta
            float* lse2 = vec + (it & 1) * 2 * SP;
            float* delta = lse2 + SP;
            {
                const float* Lg = p.lse + ((long long)b * NH + h) * S;
                for (int i = threadIdx.x - 64; i < SP; i += 128 * NSPLIT) lse2[i] = i < S ? Lg[i] * LOG2E : INFINITY;
                for (int r0 = ew * 4; r0 < SP; r0 += 16 * NSPLIT) {
                    const int r = r0 + (lane >> 3), c8 = (lane & 7) * 8;
                    float acc = 0.f;
                    if (r < S) {
                        float a[8], d[8];
                        load8_bf16(p.ctx + (row0 + r) * HID + h * HD + c8, a);
                        load8_bf16(p.dctx + (row0 + r) * HID + h * HD + c8, d);
#pragma unroll
                        for (int k = 0; k < 8; ++k) acc = fmaf(a[k], d[k], acc);
                    }
                    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
                    acc += __shfl_xor_sync(0xffffffffu, acc, 2);
                    acc += __shfl_xor_sync(0xffffffffu, acc, 4);
                    if ((lane & 7) == 0 && r < SP) delta[r] = acc;
                }
            }
            asm volatile("bar.sync 1, %0;" ::"n"(128 * NSPLIT) : "memory");   // the element-wise warps only
            const uint32_t hkey = drop_head_key(p.drop.key, b * NH + h);

            for (int u = 0; u < nu; ++u) {
                if (u * 128 + q * 32 >= S) continue;               // no key rows of this tile in this lane quarter
                const int key = u * 128 + kr;
                const bool key_ok = key < S;
                const float bias = key_ok ? (p.mask[row0 + key] != 0 ? 0.f : MASK_LOG2) : -INFINITY;
                ptx::mbar_wait(sd_full(u), itp);
                ptx::tc_fence_after();
                const uint32_t prow = static_cast<uint32_t>(kr) * 128u;
                for (int c = c_begin; c < c_end; c += 16) {
                    uint32_t rs[16], rd[16], pd[8], ds[8];
                    ptx::tmem_ld_32x16(tmem_base + lane_sel + TMB_S + c, rs);
                    ptx::tmem_ld_32x16(tmem_base + lane_sel + TMB_DP + c, rd);
                    ptx::tmem_wait_ld();
#pragma unroll
                    for (int j = 0; j < 16; j += 2) {
                        const float2 l2 = *reinterpret_cast<const float2*>(lse2 + c + j);
                        const float2 de = *reinterpret_cast<const float2*>(delta + c + j);
                        const float p0 = fast_ex2(fmaf(__uint_as_float(rs[j]), SCALE_LOG2, bias) - l2.x);
                        const float p1 = fast_ex2(fmaf(__uint_as_float(rs[j + 1]), SCALE_LOG2, bias) - l2.y);
                        float pd0 = p0, pd1 = p1, dp0 = __uint_as_float(rd[j]), dp1 = __uint_as_float(rd[j + 1]);
                        if (p.drop.thresh) {
                            // element (query c + j, key): the same counter stream as the forward (idx = query * S + key)
                            const uint32_t i0 = static_cast<uint32_t>(c + j) * static_cast<uint32_t>(S) + key;
                            const bool k0 = drop_keep(hkey, i0, p.drop.thresh);
                            const bool k1 = drop_keep(hkey, i0 + static_cast<uint32_t>(S), p.drop.thresh);
                            const float s0 = k0 ? p.drop.scale : 0.f, s1 = k1 ? p.drop.scale : 0.f;   // finite operands only
                            pd0 = p0 * s0; dp0 *= s0;
                            pd1 = p1 * s1; dp1 *= s1;
                        }
                        pd[j / 2] = pack_bf16(pd0, pd1);
                        // dS without the 1/sqrt(64): a power of two commutes with the bf16 rounding, so it is applied
                        // once per OUTPUT element in the dK / dQ epilogues instead of once per score here
                        ds[j / 2] = pack_bf16(p0 * (dp0 - de.x), p1 * (dp1 - de.y));
                    }
#pragma unroll
                    for (int g = 0; g < 2; ++g) {
                        const uint32_t q0 = c + 8 * g;             // 8 queries = one 16-byte unit of the swizzled row
                        const uint32_t off = (q0 >> 6) * P_CHUNK_BYTES + prow + ((((q0 & 63u) >> 3) ^ sw) << 4);
                        ptx::st_shared_v4(sPD + off, pd[4 * g], pd[4 * g + 1], pd[4 * g + 2], pd[4 * g + 3]);
                        ptx::st_shared_v4(sDS + off, ds[4 * g], ds[4 * g + 1], ds[4 * g + 2], ds[4 * g + 3]);
                    }
                }
                ptx::fence_proxy_async();
                ptx::tc_fence_before();
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive(pds_full(u));

                // this key row's [dV_u | dK_u] = TMEM columns [0,128): part -> a 128 / NSPLIT-column slice -> dqkv
                ptx::mbar_wait(kv_full(u), itp);
                ptx::tc_fence_after();
                constexpr int WKV = 128 / NSPLIT;
                const int ckv = part * WKV;
                bf16* dst = p.dqkv + (row0 + key) * QKV_LD + (ckv < HD ? 2 * HID : HID) + h * HD + (ckv & (HD - 1));
                store_cols<WKV>(tmem_base + lane_sel + TMB_DV + ckv, dst, key_ok, ckv < HD ? 1.f : 0.125f);
                ptx::tc_fence_before();
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive(s_empty(u));
            }

            // dQ: query rows in the lanes; part = which 64 / NSPLIT of the 64 columns
            if (q * 32 < S) {
                ptx::mbar_wait(dq_full, itp);
                ptx::tc_fence_after();
                constexpr int WQ = HD / NSPLIT;
                store_cols<WQ>(tmem_base + lane_sel + TMB_DQ + part * WQ,
                               p.dqkv + (row0 + kr) * QKV_LD + h * HD + part * WQ, kr < S, 0.125f);
                if (nu == 2 && q == 0)
                    store_cols<WQ>(tmem_base + TMB_DQ + 64u + part * WQ,
                                   p.dqkv + (row0 + 128 + lane) * QKV_LD + h * HD + part * WQ, 128 + lane < S, 0.125f);
                ptx::tc_fence_before();
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive(dq_empty);
            }
        }
    }

    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

// Both kernels allocate all 512 TMEM columns, so two of their CTAs must never share an SM (the second would sit in
// tcgen05.alloc until the first one's whole persistent loop is over): short sequences, whose tiles are small, still
// ask for more than half of the SM's shared memory.
inline size_t exclusive_smem(uint32_t need) { return need > 120u * 1024u ? need : 120u * 1024u; }

std::atomic<int> g_tc_enabled{-1};      // -1: not decided yet (environment), 0 / 1

}  // namespace

bool attn_tc_enabled() {
    int v = g_tc_enabled.load(std::memory_order_relaxed);
    if (v < 0) {
        const char* e = getenv("UC2_ATTN_TCGEN05");
        v = (e && e[0] == '1') ? 1 : 0;
        g_tc_enabled.store(v, std::memory_order_relaxed);
    }
    return v == 1;
}

}  // namespace uc2

using namespace uc2;

extern "C" UC2_API int uc2_attention_tc_enable(int on) {
    const int prev = attn_tc_enabled() ? 1 : 0;
    g_tc_enabled.store(on ? 1 : 0, std::memory_order_relaxed);
    return prev;
}

extern "C" UC2_API int uc2_attention_fwd_tc(const void* qkv, const long long* attn_mask, void* ctx, float* lse, int B,
                                            int S, unsigned int drop_key, unsigned int drop_thresh, float drop_scale,
                                            void* stream) {
    if (int rc = require_sm100()) return rc;
    UC2_REQUIRE(qkv && attn_mask && ctx && lse, UC2_ERR_ARG, "attention_fwd_tc: null pointer");
    UC2_REQUIRE(B > 0 && S > 0, UC2_ERR_ARG, "attention_fwd_tc: bad shape B=%d S=%d", B, S);
    UC2_REQUIRE(drop_thresh < 65536u, UC2_ERR_ARG, "attention_fwd_tc: drop_thresh must be < 65536");
    const int SP = (S + 15) / 16 * 16;
    UC2_REQUIRE(SP <= TC_MAX_SP, UC2_ERR_UNSUPPORTED, "attention_fwd_tc: S=%d > %d", S, TC_MAX_SP);
    UC2_REQUIRE(aligned16(qkv) && (reinterpret_cast<uintptr_t>(ctx) & 31) == 0, UC2_ERR_ARG,
                "attention_fwd_tc: qkv must be 16-byte and ctx 32-byte aligned");
    // qkv as a 2-D bf16 tensor [B*S][2304]; one box = the SP x 64 tile of one head's Q, K or V (rows past B*S are
    // zero-filled, rows past S inside the box belong to the next sample and are masked / never stored)
    CUtensorMap tmap;
    if (int rc = make_tmap(&tmap, qkv, (long long)B * S, QKV_LD, QKV_LD, SP)) return rc;
    const TcSmem L = tc_smem(S, SP);
    static std::once_flag once;
    static cudaError_t attr_err = cudaSuccess;
    std::call_once(once, [] {
        attr_err = cudaFuncSetAttribute(attention_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        static_cast<int>(tc_smem(TC_MAX_SP, TC_MAX_SP).total));
    });
    UC2_REQUIRE(attr_err == cudaSuccess, UC2_ERR_CUDA, "attention_fwd_tc: cudaFuncSetAttribute failed: %s",
                cudaGetErrorString(attr_err));
    TcParams p;
    p.mask = attn_mask;
    p.ctx = static_cast<bf16*>(ctx);
    p.lse = lse;
    p.B = B; p.S = S; p.SP = SP; p.items = B * NH;
    p.drop = DropCfg{drop_key, drop_thresh, drop_scale};
    const int grid = p.items < num_sms() ? p.items : num_sms();
    ProfScope prof((cudaStream_t)stream, 1, 4.0 * B * NH * (double)S * S * HD);
    const cudaError_t e = launch_pdl(attention_fwd_tc_kernel, dim3(grid), dim3(TC_THREADS), exclusive_smem(L.total),
                                     (cudaStream_t)stream, 1, tmap, p);
    UC2_REQUIRE(e == cudaSuccess, UC2_ERR_CUDA, "attention_fwd_tc launch failed: %s", cudaGetErrorString(e));
    return check_last("attention_fwd_tc_kernel");
}

extern "C" UC2_API int uc2_attention_bwd_tc(const void* qkv, const long long* attn_mask, const void* ctx,
                                            const void* dctx, const float* lse, void* dqkv, int B, int S,
                                            unsigned int drop_key, unsigned int drop_thresh, float drop_scale,
                                            void* stream) {
    if (int rc = require_sm100()) return rc;
    UC2_REQUIRE(qkv && attn_mask && ctx && dctx && lse && dqkv, UC2_ERR_ARG, "attention_bwd_tc: null pointer");
    UC2_REQUIRE(B > 0 && S > 0, UC2_ERR_ARG, "attention_bwd_tc: bad shape B=%d S=%d", B, S);
    UC2_REQUIRE(drop_thresh < 65536u, UC2_ERR_ARG, "attention_bwd_tc: drop_thresh must be < 65536");
    const int SP = (S + 15) / 16 * 16;
    UC2_REQUIRE(SP <= TC_MAX_SP, UC2_ERR_UNSUPPORTED, "attention_bwd_tc: S=%d > %d", S, TC_MAX_SP);
    UC2_REQUIRE(aligned16(qkv) && aligned16(ctx) && aligned16(dctx) && (reinterpret_cast<uintptr_t>(dqkv) & 31) == 0,
                UC2_ERR_ARG, "attention_bwd_tc: qkv / ctx / dctx must be 16-byte and dqkv 32-byte aligned");
    CUtensorMap tq, tdo;
    if (int rc = make_tmap(&tq, qkv, (long long)B * S, QKV_LD, QKV_LD, SP)) return rc;
    if (int rc = make_tmap(&tdo, dctx, (long long)B * S, HID, HID, SP)) return rc;
    const TcBwdSmem L = tc_bwd_smem(S, SP);
    static std::once_flag once;
    static cudaError_t attr_err = cudaSuccess;
    static int nsplit = 2;
    std::call_once(once, [] {
        const char* e = getenv("UC2_ATTN_TC_BWD_SPLIT");          // tuning knob: 2 (default) or 4 warps per lane quarter
        if (e && e[0] == '4') nsplit = 4;
        const int bytes = static_cast<int>(tc_bwd_smem(TC_MAX_SP, TC_MAX_SP).total);
        attr_err = cudaFuncSetAttribute(attention_bwd_tc_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
        if (attr_err == cudaSuccess)
            attr_err = cudaFuncSetAttribute(attention_bwd_tc_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    });
    UC2_REQUIRE(attr_err == cudaSuccess, UC2_ERR_CUDA, "attention_bwd_tc: cudaFuncSetAttribute failed: %s",
                cudaGetErrorString(attr_err));
    TcBwdParams p;
    p.mask = attn_mask;
    p.ctx = static_cast<const bf16*>(ctx);
    p.dctx = static_cast<const bf16*>(dctx);
    p.lse = lse;
    p.dqkv = static_cast<bf16*>(dqkv);
    p.B = B; p.S = S; p.SP = SP; p.items = B * NH;
    p.drop = DropCfg{drop_key, drop_thresh, drop_scale};
    const int grid = p.items < num_sms() ? p.items : num_sms();
    ProfScope prof((cudaStream_t)stream, 1, 10.0 * B * NH * (double)S * S * HD);
    const cudaError_t e =
        nsplit == 4 ? launch_pdl(attention_bwd_tc_kernel<4>, dim3(grid), dim3(bwd_threads(4)), exclusive_smem(L.total),
                                 (cudaStream_t)stream, 1, tq, tdo, p)
                    : launch_pdl(attention_bwd_tc_kernel<2>, dim3(grid), dim3(bwd_threads(2)), exclusive_smem(L.total),
                                 (cudaStream_t)stream, 1, tq, tdo, p);
    UC2_REQUIRE(e == cudaSuccess, UC2_ERR_CUDA, "attention_bwd_tc launch failed: %s", cudaGetErrorString(e));
    return check_last("attention_bwd_tc_kernel");
}
