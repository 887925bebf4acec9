// Flash-style key-padding-masked self-attention for 12 heads x 64 (forward + backward), S <= 512.
// Replaces BertSelfAttention.forward model/layer.py:80-100 (scores, additive mask, softmax, P.V,
// head merge) without materialising the two [B,12,S,S] tensors of the reference.
// Warp-level mma.sync m16n8k16 (bf16 in, fp32 accumulate); attention is 3.4% of the encoder FLOPs at S=160
// (SURVEY 8d).  Two generations of kernels live here:
//   * S <= 256 (every BASELINE shape): one CTA per (batch, head) with Q, K, V (and dO) of that head resident in
//     shared memory ONCE, one warp per 16 rows, 32-wide inner chunks (so S = 160 costs 160, not 192, in both
//     dimensions).  The backward kernel is persistent and double-buffers the next head's tiles with cp.async
//     while it computes dQ (phase A, warp = 16 query rows) and dK/dV (phase B, warp = 16 key rows).
//   * S <= 512: the tiled kernels (64-row query / key tiles per CTA) kept as the general path.
#include "common.cuh"
#include "ptx.cuh"

namespace uc2 {
namespace {

constexpr int HD = 64;            // head dim
constexpr int NH = 12;
constexpr int QKV_LD = 3 * HID;   // 2304
constexpr int ATT_THREADS = 128;  // 4 warps x 16 rows
constexpr int TILE = 64;
constexpr float LOG2E = 1.4426950408889634f;
constexpr float SCALE_LOG2 = 0.125f * LOG2E;   // 1/sqrt(64) folded into the exp2 domain
constexpr float MASK_LOG2 = -10000.0f * LOG2E;

// swizzled [rows][64] bf16 tile: 16-byte chunk c of row r lives at chunk (c ^ (r & 7))
__device__ __forceinline__ uint32_t tile_off(int row, int chunk) { return row * 128 + ((chunk ^ (row & 7)) << 4); }

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool valid) {
    const int sz = valid ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
    asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
}

// rows [row0, row0+nrows) of a [*, ld] bf16 matrix (64 columns starting at col0) -> swizzled tile; rows >= limit zero
__device__ __forceinline__ void load_tile(uint32_t smem, const bf16* g, long long ld, long long grow0, int col0,
                                          int nrows, int valid_rows) {
    for (int i = threadIdx.x; i < nrows * 8; i += ATT_THREADS) {
        const int r = i >> 3, c = i & 7;
        const bool ok = r < valid_rows;
        const bf16* src = g + (grow0 + (ok ? r : 0)) * ld + col0 + c * 8;
        cp_async16(smem + tile_off(r, c), src, ok);
    }
}

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t* r) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t* r) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void mma16816(float* c, const uint32_t* a, uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
        "{%0, %1, %2, %3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// A fragments (16 rows x 64 k) of the rows [row0, row0+16) of a tile
__device__ __forceinline__ void load_a_frags(uint32_t tile, int row0, int lane, uint32_t f[4][4]) {
#pragma unroll
    for (int kk = 0; kk < 4; ++kk)
        ldsm_x4(tile + tile_off(row0 + (lane & 15), kk * 2 + (lane >> 4)), f[kk]);
}

// C[16 x 64] += A[16 x 64(k)] * B^T where B tile rows are the n index ([n][k] row-major): rows n0..n0+63
__device__ __forceinline__ void mma_nk(float c[8][4], const uint32_t a[4][4], uint32_t tile, int n0, int lane) {
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            uint32_t r[4];
            ldsm_x4(tile + tile_off(n0 + p * 16 + (lane & 7) + ((lane >> 4) << 3), kk * 2 + ((lane >> 3) & 1)), r);
            mma16816(c[2 * p], a[kk], r[0], r[1]);
            mma16816(c[2 * p + 1], a[kk], r[2], r[3]);
        }
    }
}

// C[16 x 64(n)] += A[16 x 64(k)] * B where B tile rows are the k index ([k][n] row-major): rows k0..k0+63.
// A comes from fp32 accumulator fragments p[8][4] (C layout of a 16x64 product) converted to bf16.
__device__ __forceinline__ void mma_kn_from_c(float c[8][4], const float p[8][4], uint32_t tile, int k0, int lane) {
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
        uint32_t a[4];
        a[0] = pack_bf16(p[2 * kk][0], p[2 * kk][1]);
        a[1] = pack_bf16(p[2 * kk][2], p[2 * kk][3]);
        a[2] = pack_bf16(p[2 * kk + 1][0], p[2 * kk + 1][1]);
        a[3] = pack_bf16(p[2 * kk + 1][2], p[2 * kk + 1][3]);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            uint32_t r[4];
            ldsm_x4_t(tile + tile_off(k0 + kk * 16 + (lane & 7) + (((lane >> 3) & 1) << 3), q * 2 + (lane >> 4)), r);
            mma16816(c[2 * q], a, r[0], r[1]);
            mma16816(c[2 * q + 1], a, r[2], r[3]);
        }
    }
}

__device__ __forceinline__ float quad_max(float v) {
    v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
    return fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
}
__device__ __forceinline__ float quad_sum(float v) {
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    return v + __shfl_xor_sync(0xffffffffu, v, 2);
}

// write a 16x64 C-layout fragment as bf16 rows of `out` (leading dim ld), rows >= valid skipped
__device__ __forceinline__ void store_c_bf16(bf16* out, long long ld, long long grow0, int col0, int lane,
                                             const float c[8][4], int valid_rows, int row0_local) {
    const int g = lane >> 2, t = lane & 3;
#pragma unroll
    for (int n = 0; n < 8; ++n) {
        const int col = col0 + n * 8 + 2 * t;
        if (row0_local + g < valid_rows)
            *reinterpret_cast<uint32_t*>(out + (grow0 + g) * ld + col) = pack_bf16(c[n][0], c[n][1]);
        if (row0_local + g + 8 < valid_rows)
            *reinterpret_cast<uint32_t*>(out + (grow0 + g + 8) * ld + col) = pack_bf16(c[n][2], c[n][3]);
    }
}

// shared layout helpers -----------------------------------------------------------------------
// forward / dQ:  [Q 64][dO 64 (dQ only)][K S_pad][V S_pad][maskbias S_pad floats]
// dKV:           [K 64][V 64][Q S_pad][dO S_pad][lse S_pad f][D S_pad f]

__global__ void __launch_bounds__(ATT_THREADS)
attention_fwd_kernel(const bf16* __restrict__ qkv, const long long* __restrict__ mask, bf16* __restrict__ ctx,
                     float* __restrict__ lse, int S, int S_pad) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t sQ = static_cast<uint32_t>(__cvta_generic_to_shared(smem));
    const uint32_t sK = sQ + TILE * 128;
    const uint32_t sV = sK + S_pad * 128;
    float* mbias = reinterpret_cast<float*>(smem + TILE * 128 + 2 * S_pad * 128);
    const long long base = (long long)b * S;
    const int q0 = qt * TILE;
    load_tile(sQ, qkv, QKV_LD, base + q0, h * HD, TILE, min(TILE, S - q0));
    load_tile(sK, qkv, QKV_LD, base, HID + h * HD, S_pad, S);
    load_tile(sV, qkv, QKV_LD, base, 2 * HID + h * HD, S_pad, S);
    for (int i = threadIdx.x; i < S_pad; i += ATT_THREADS)
        mbias[i] = i < S ? (mask[base + i] != 0 ? 0.f : MASK_LOG2) : -INFINITY;
    cp_async_wait_all();
    __syncthreads();
    if (q0 + warp * 16 >= S) return;
    uint32_t qf[4][4];
    load_a_frags(sQ, warp * 16, lane, qf);
    float o[8][4];
#pragma unroll
    for (int n = 0; n < 8; ++n) o[n][0] = o[n][1] = o[n][2] = o[n][3] = 0.f;
    float m_lo = -INFINITY, m_hi = -INFINITY, l_lo = 0.f, l_hi = 0.f;
    const int t = lane & 3;
    for (int kc = 0; kc < S_pad; kc += TILE) {
        float s[8][4];
#pragma unroll
        for (int n = 0; n < 8; ++n) s[n][0] = s[n][1] = s[n][2] = s[n][3] = 0.f;
        mma_nk(s, qf, sK, kc, lane);
        float mx_lo = m_lo, mx_hi = m_hi;
#pragma unroll
        for (int n = 0; n < 8; ++n) {
            const float b0 = mbias[kc + n * 8 + 2 * t], b1 = mbias[kc + n * 8 + 2 * t + 1];
            s[n][0] = s[n][0] * SCALE_LOG2 + b0; s[n][1] = s[n][1] * SCALE_LOG2 + b1;
            s[n][2] = s[n][2] * SCALE_LOG2 + b0; s[n][3] = s[n][3] * SCALE_LOG2 + b1;
            mx_lo = fmaxf(mx_lo, fmaxf(s[n][0], s[n][1]));
            mx_hi = fmaxf(mx_hi, fmaxf(s[n][2], s[n][3]));
        }
        mx_lo = quad_max(mx_lo); mx_hi = quad_max(mx_hi);
        const float c_lo = exp2f(m_lo - mx_lo), c_hi = exp2f(m_hi - mx_hi);
        m_lo = mx_lo; m_hi = mx_hi;
        l_lo *= c_lo; l_hi *= c_hi;
#pragma unroll
        for (int n = 0; n < 8; ++n) {
            s[n][0] = exp2f(s[n][0] - m_lo); s[n][1] = exp2f(s[n][1] - m_lo);
            s[n][2] = exp2f(s[n][2] - m_hi); s[n][3] = exp2f(s[n][3] - m_hi);
            l_lo += s[n][0] + s[n][1]; l_hi += s[n][2] + s[n][3];
            o[n][0] *= c_lo; o[n][1] *= c_lo; o[n][2] *= c_hi; o[n][3] *= c_hi;
        }
        mma_kn_from_c(o, s, sV, kc, lane);
    }
    l_lo = quad_sum(l_lo); l_hi = quad_sum(l_hi);
    const float i_lo = 1.f / l_lo, i_hi = 1.f / l_hi;
#pragma unroll
    for (int n = 0; n < 8; ++n) { o[n][0] *= i_lo; o[n][1] *= i_lo; o[n][2] *= i_hi; o[n][3] *= i_hi; }
    const int rl = q0 + warp * 16;
    store_c_bf16(ctx, HID, base + rl, h * HD, lane, o, S, rl);
    if (t == 0) {
        const int g = lane >> 2;
        float* L = lse + ((long long)b * NH + h) * S;
        if (rl + g < S) L[rl + g] = (m_lo + log2f(l_lo)) * (1.f / LOG2E);
        if (rl + g + 8 < S) L[rl + g + 8] = (m_hi + log2f(l_hi)) * (1.f / LOG2E);
    }
}

// D[b,h,q] = sum_d dO[q,h,d] * O[q,h,d]
__global__ void __launch_bounds__(256)
attention_delta_kernel(const bf16* __restrict__ ctx, const bf16* __restrict__ dctx, float* __restrict__ delta,
                       int B, int S) {
    const int lane = threadIdx.x & 31;
    const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    griddep_sync();
    if (row >= (long long)B * S) return;
    const int b = (int)(row / S), q = (int)(row % S);
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        float a[8], d[8];
        const int col = i * 256 + lane * 8;
        load8_bf16(ctx + row * HID + col, a);
        load8_bf16(dctx + row * HID + col, d);
        float s = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) s += a[k] * d[k];
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        s += __shfl_xor_sync(0xffffffffu, s, 4);
        if ((lane & 7) == 0) delta[((long long)b * NH + (col >> 6)) * S + q] = s;
    }
}

__global__ void __launch_bounds__(ATT_THREADS)
attention_bwd_dq_kernel(const bf16* __restrict__ qkv, const long long* __restrict__ mask,
                        const bf16* __restrict__ dctx, const float* __restrict__ lse,
                        const float* __restrict__ delta, bf16* __restrict__ dqkv, int S, int S_pad) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t sQ = static_cast<uint32_t>(__cvta_generic_to_shared(smem));
    const uint32_t sdO = sQ + TILE * 128;
    const uint32_t sK = sdO + TILE * 128;
    const uint32_t sV = sK + S_pad * 128;
    float* mbias = reinterpret_cast<float*>(smem + 2 * TILE * 128 + 2 * S_pad * 128);
    const long long base = (long long)b * S;
    const int q0 = qt * TILE;
    const int vq = min(TILE, S - q0);
    load_tile(sQ, qkv, QKV_LD, base + q0, h * HD, TILE, vq);
    load_tile(sdO, dctx, HID, base + q0, h * HD, TILE, vq);
    load_tile(sK, qkv, QKV_LD, base, HID + h * HD, S_pad, S);
    load_tile(sV, qkv, QKV_LD, base, 2 * HID + h * HD, S_pad, S);
    for (int i = threadIdx.x; i < S_pad; i += ATT_THREADS)
        mbias[i] = i < S ? (mask[base + i] != 0 ? 0.f : MASK_LOG2) : -INFINITY;
    cp_async_wait_all();
    __syncthreads();
    const int rl = q0 + warp * 16;
    if (rl >= S) return;
    uint32_t qf[4][4], dof[4][4];
    load_a_frags(sQ, warp * 16, lane, qf);
    load_a_frags(sdO, warp * 16, lane, dof);
    const int g = lane >> 2, t = lane & 3;
    const float* L = lse + ((long long)b * NH + h) * S;
    const float* Dl = delta + ((long long)b * NH + h) * S;
    const float lse_lo = rl + g < S ? L[rl + g] * LOG2E : INFINITY;
    const float lse_hi = rl + g + 8 < S ? L[rl + g + 8] * LOG2E : INFINITY;
    const float d_lo = rl + g < S ? Dl[rl + g] : 0.f;
    const float d_hi = rl + g + 8 < S ? Dl[rl + g + 8] : 0.f;
    float dq[8][4];
#pragma unroll
    for (int n = 0; n < 8; ++n) dq[n][0] = dq[n][1] = dq[n][2] = dq[n][3] = 0.f;
    for (int kc = 0; kc < S_pad; kc += TILE) {
        float s[8][4], dp[8][4];
#pragma unroll
        for (int n = 0; n < 8; ++n) {
            s[n][0] = s[n][1] = s[n][2] = s[n][3] = 0.f;
            dp[n][0] = dp[n][1] = dp[n][2] = dp[n][3] = 0.f;
        }
        mma_nk(s, qf, sK, kc, lane);
        mma_nk(dp, dof, sV, kc, lane);
#pragma unroll
        for (int n = 0; n < 8; ++n) {
            const float b0 = mbias[kc + n * 8 + 2 * t], b1 = mbias[kc + n * 8 + 2 * t + 1];
            const float p0 = exp2f(s[n][0] * SCALE_LOG2 + b0 - lse_lo), p1 = exp2f(s[n][1] * SCALE_LOG2 + b1 - lse_lo);
            const float p2 = exp2f(s[n][2] * SCALE_LOG2 + b0 - lse_hi), p3 = exp2f(s[n][3] * SCALE_LOG2 + b1 - lse_hi);
            s[n][0] = p0 * (dp[n][0] - d_lo) * 0.125f; s[n][1] = p1 * (dp[n][1] - d_lo) * 0.125f;
            s[n][2] = p2 * (dp[n][2] - d_hi) * 0.125f; s[n][3] = p3 * (dp[n][3] - d_hi) * 0.125f;
        }
        mma_kn_from_c(dq, s, sK, kc, lane);
    }
    store_c_bf16(dqkv, QKV_LD, base + rl, h * HD, lane, dq, S, rl);
}

__global__ void __launch_bounds__(ATT_THREADS)
attention_bwd_dkv_kernel(const bf16* __restrict__ qkv, const long long* __restrict__ mask,
                         const bf16* __restrict__ dctx, const float* __restrict__ lse,
                         const float* __restrict__ delta, bf16* __restrict__ dqkv, int S, int S_pad) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int kt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t sK = static_cast<uint32_t>(__cvta_generic_to_shared(smem));
    const uint32_t sV = sK + TILE * 128;
    const uint32_t sQ = sV + TILE * 128;
    const uint32_t sdO = sQ + S_pad * 128;
    float* s_lse = reinterpret_cast<float*>(smem + 2 * TILE * 128 + 2 * S_pad * 128);
    float* s_del = s_lse + S_pad;
    const long long base = (long long)b * S;
    const int k0 = kt * TILE;
    const int vk = min(TILE, S - k0);
    load_tile(sK, qkv, QKV_LD, base + k0, HID + h * HD, TILE, vk);
    load_tile(sV, qkv, QKV_LD, base + k0, 2 * HID + h * HD, TILE, vk);
    load_tile(sQ, qkv, QKV_LD, base, h * HD, S_pad, S);
    load_tile(sdO, dctx, HID, base, h * HD, S_pad, S);
    const float* L = lse + ((long long)b * NH + h) * S;
    const float* Dl = delta + ((long long)b * NH + h) * S;
    for (int i = threadIdx.x; i < S_pad; i += ATT_THREADS) {
        s_lse[i] = i < S ? L[i] * LOG2E : INFINITY;      // +inf: padded query rows contribute exp2(-inf) = 0
        s_del[i] = i < S ? Dl[i] : 0.f;
    }
    cp_async_wait_all();
    __syncthreads();
    const int rl = k0 + warp * 16;
    if (rl >= S) return;
    uint32_t kf[4][4], vf[4][4];
    load_a_frags(sK, warp * 16, lane, kf);
    load_a_frags(sV, warp * 16, lane, vf);
    const int g = lane >> 2, t = lane & 3;
    const float mb_lo = rl + g < S ? (mask[base + rl + g] != 0 ? 0.f : MASK_LOG2) : -INFINITY;
    const float mb_hi = rl + g + 8 < S ? (mask[base + rl + g + 8] != 0 ? 0.f : MASK_LOG2) : -INFINITY;
    float dk[8][4], dv[8][4];
#pragma unroll
    for (int n = 0; n < 8; ++n) {
        dk[n][0] = dk[n][1] = dk[n][2] = dk[n][3] = 0.f;
        dv[n][0] = dv[n][1] = dv[n][2] = dv[n][3] = 0.f;
    }
    for (int qc = 0; qc < S_pad; qc += TILE) {
        float st[8][4], dpt[8][4];
#pragma unroll
        for (int n = 0; n < 8; ++n) {
            st[n][0] = st[n][1] = st[n][2] = st[n][3] = 0.f;
            dpt[n][0] = dpt[n][1] = dpt[n][2] = dpt[n][3] = 0.f;
        }
        mma_nk(st, kf, sQ, qc, lane);        // S^T[key, q]
        mma_nk(dpt, vf, sdO, qc, lane);      // dP^T[key, q]
        float pt[8][4];
#pragma unroll
        for (int n = 0; n < 8; ++n) {
            const int c = qc + n * 8 + 2 * t;
            const float l0 = s_lse[c], l1 = s_lse[c + 1], e0 = s_del[c], e1 = s_del[c + 1];
            pt[n][0] = exp2f(st[n][0] * SCALE_LOG2 + mb_lo - l0); pt[n][1] = exp2f(st[n][1] * SCALE_LOG2 + mb_lo - l1);
            pt[n][2] = exp2f(st[n][2] * SCALE_LOG2 + mb_hi - l0); pt[n][3] = exp2f(st[n][3] * SCALE_LOG2 + mb_hi - l1);
            st[n][0] = pt[n][0] * (dpt[n][0] - e0) * 0.125f; st[n][1] = pt[n][1] * (dpt[n][1] - e1) * 0.125f;
            st[n][2] = pt[n][2] * (dpt[n][2] - e0) * 0.125f; st[n][3] = pt[n][3] * (dpt[n][3] - e1) * 0.125f;
        }
        mma_kn_from_c(dv, pt, sdO, qc, lane);
        mma_kn_from_c(dk, st, sQ, qc, lane);
    }
    store_c_bf16(dqkv, QKV_LD, base + rl, HID + h * HD, lane, dk, S, rl);
    store_c_bf16(dqkv, QKV_LD, base + rl, 2 * HID + h * HD, lane, dv, S, rl);
}

// ---------------------------------------------------------------------------------------------------------
// per-(batch, head) kernels for S <= 256
// ---------------------------------------------------------------------------------------------------------
constexpr int KC = 32;            // inner chunk (keys in phase A / forward, queries in phase B)
constexpr int BH_MAX_WARPS = 10;      // S = 160 -> one 16-row tile per warp

__device__ __forceinline__ float ex2a(float x) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// rows of a [*, ld] bf16 matrix (64 columns starting at col0) -> swizzled tile, any thread count
__device__ __forceinline__ void load_tile_n(uint32_t smem, const bf16* g, long long ld, long long grow0, int col0,
                                            int nrows, int valid_rows, int nthreads) {
    for (int i = threadIdx.x; i < nrows * 8; i += nthreads) {
        const int r = i >> 3, c = i & 7;
        const bool ok = r < valid_rows;
        const bf16* src = g + (grow0 + (ok ? r : 0)) * ld + col0 + c * 8;
        cp_async16(smem + tile_off(r, c), src, ok);
    }
}

// C[16 x 16 NP] += A[16 x 64(k)] * B^T, B tile rows n0.. are the n index ([n][k] row-major)
template <int NP>
__device__ __forceinline__ void mma_nk_t(float (*c)[4], const uint32_t a[4][4], uint32_t tile, int n0, int lane) {
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
        for (int p = 0; p < NP; ++p) {
            uint32_t r[4];
            ldsm_x4(tile + tile_off(n0 + p * 16 + (lane & 7) + ((lane >> 4) << 3), kk * 2 + ((lane >> 3) & 1)), r);
            mma16816(c[2 * p], a[kk], r[0], r[1]);
            mma16816(c[2 * p + 1], a[kk], r[2], r[3]);
        }
    }
}

// C[16 x 64(n)] += P[16 x 16 NK] * B, P = fp32 C-layout fragments, B tile rows k0.. are the k index ([k][n])
template <int NK>
__device__ __forceinline__ void mma_kn_t(float c[8][4], const float (*p)[4], uint32_t tile, int k0, int lane) {
#pragma unroll
    for (int kk = 0; kk < NK; ++kk) {
        uint32_t a[4];
        a[0] = pack_bf16(p[2 * kk][0], p[2 * kk][1]);
        a[1] = pack_bf16(p[2 * kk][2], p[2 * kk][3]);
        a[2] = pack_bf16(p[2 * kk + 1][0], p[2 * kk + 1][1]);
        a[3] = pack_bf16(p[2 * kk + 1][2], p[2 * kk + 1][3]);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            uint32_t r[4];
            ldsm_x4_t(tile + tile_off(k0 + kk * 16 + (lane & 7) + (((lane >> 3) & 1) << 3), q * 2 + (lane >> 4)), r);
            mma16816(c[2 * q], a, r[0], r[1]);
            mma16816(c[2 * q + 1], a, r[2], r[3]);
        }
    }
}

// grid (NH, B); block = NW warps; shared: [Q SP][K SP][V SP][mask bias SP floats], SP = S rounded up to 32
__global__ void __launch_bounds__(BH_MAX_WARPS * 32, 2)
attention_fwd_bh_kernel(const bf16* __restrict__ qkv, const long long* __restrict__ mask, bf16* __restrict__ ctx,
                        float* __restrict__ lse, int S, int SP, DropCfg drop) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int h = blockIdx.x, b = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    const uint32_t sQ = static_cast<uint32_t>(__cvta_generic_to_shared(smem));
    const uint32_t sK = sQ + SP * 128;
    const uint32_t sV = sK + SP * 128;
    float* mbias = reinterpret_cast<float*>(smem + 3 * SP * 128);
    const long long base = (long long)b * S;
    griddep_sync();
    load_tile_n(sQ, qkv, QKV_LD, base, h * HD, SP, S, blockDim.x);
    load_tile_n(sK, qkv, QKV_LD, base, HID + h * HD, SP, S, blockDim.x);
    load_tile_n(sV, qkv, QKV_LD, base, 2 * HID + h * HD, SP, S, blockDim.x);
    for (int i = threadIdx.x; i < SP; i += blockDim.x)
        mbias[i] = i < S ? (mask[base + i] != 0 ? 0.f : MASK_LOG2) : -INFINITY;
    cp_async_wait_all();
    __syncthreads();
    const int t = lane & 3, g = lane >> 2;
    const uint32_t hkey = drop_head_key(drop.key, b * NH + h);
    for (int q0 = warp * 16; q0 < S; q0 += nw * 16) {
        uint32_t qf[4][4];
        load_a_frags(sQ, q0, lane, qf);
        float o[8][4];
#pragma unroll
        for (int n = 0; n < 8; ++n) o[n][0] = o[n][1] = o[n][2] = o[n][3] = 0.f;
        float m_lo = -INFINITY, m_hi = -INFINITY, l_lo = 0.f, l_hi = 0.f;
        for (int kc = 0; kc < SP; kc += KC) {
            float sc[4][4];
#pragma unroll
            for (int n = 0; n < 4; ++n) sc[n][0] = sc[n][1] = sc[n][2] = sc[n][3] = 0.f;
            mma_nk_t<2>(sc, qf, sK, kc, lane);
            float mx_lo = m_lo, mx_hi = m_hi;
#pragma unroll
            for (int n = 0; n < 4; ++n) {
                const float2 bb = *reinterpret_cast<const float2*>(mbias + kc + n * 8 + 2 * t);
                sc[n][0] = fmaf(sc[n][0], SCALE_LOG2, bb.x); sc[n][1] = fmaf(sc[n][1], SCALE_LOG2, bb.y);
                sc[n][2] = fmaf(sc[n][2], SCALE_LOG2, bb.x); sc[n][3] = fmaf(sc[n][3], SCALE_LOG2, bb.y);
                mx_lo = fmaxf(mx_lo, fmaxf(sc[n][0], sc[n][1]));
                mx_hi = fmaxf(mx_hi, fmaxf(sc[n][2], sc[n][3]));
            }
            mx_lo = quad_max(mx_lo); mx_hi = quad_max(mx_hi);
            const float c_lo = ex2a(m_lo - mx_lo), c_hi = ex2a(m_hi - mx_hi);
            m_lo = mx_lo; m_hi = mx_hi;
            l_lo *= c_lo; l_hi *= c_hi;
#pragma unroll
            for (int n = 0; n < 4; ++n) {
                sc[n][0] = ex2a(sc[n][0] - m_lo); sc[n][1] = ex2a(sc[n][1] - m_lo);
                sc[n][2] = ex2a(sc[n][2] - m_hi); sc[n][3] = ex2a(sc[n][3] - m_hi);
                l_lo += sc[n][0] + sc[n][1]; l_hi += sc[n][2] + sc[n][3];
            }
#pragma unroll
            for (int n = 0; n < 8; ++n) { o[n][0] *= c_lo; o[n][1] *= c_lo; o[n][2] *= c_hi; o[n][3] *= c_hi; }
            if (drop.thresh) {
                // attention-probability dropout (layer.py:94): the normaliser keeps every key, the dropped and
                // rescaled probabilities only enter P.V
#pragma unroll
                for (int n = 0; n < 4; ++n)
#pragma unroll
                    for (int e = 0; e < 4; e += 2) {
                        const uint32_t idx = (uint32_t)(q0 + g + (e >> 1) * 8) * (uint32_t)S + kc + n * 8 + 2 * t;
                        bool k0, k1;
                        drop_keep2(hkey, idx, drop.thresh, k0, k1);
                        sc[n][e] = k0 ? sc[n][e] * drop.scale : 0.f;
                        sc[n][e + 1] = k1 ? sc[n][e + 1] * drop.scale : 0.f;
                    }
            }
            mma_kn_t<2>(o, sc, sV, kc, lane);
        }
        l_lo = quad_sum(l_lo); l_hi = quad_sum(l_hi);
        const float i_lo = 1.f / l_lo, i_hi = 1.f / l_hi;
#pragma unroll
        for (int n = 0; n < 8; ++n) { o[n][0] *= i_lo; o[n][1] *= i_lo; o[n][2] *= i_hi; o[n][3] *= i_hi; }
        store_c_bf16(ctx, HID, base + q0, h * HD, lane, o, S, q0);
        if (t == 0) {
            float* L = lse + ((long long)b * NH + h) * S;
            if (q0 + g < S) L[q0 + g] = (m_lo + log2f(l_lo)) * (1.f / LOG2E);
            if (q0 + g + 8 < S) L[q0 + g + 8] = (m_hi + log2f(l_hi)) * (1.f / LOG2E);
        }
    }
}

// Persistent backward: grid = min(B * NH, #SMs) CTAs of NW warps looping over (batch, head) items.
// shared per buffer: [Q SP][K SP][V SP][dO SP] tiles; then per CTA: lse2[SP], delta[SP], mbias[SP] floats.
__global__ void __launch_bounds__(BH_MAX_WARPS * 32, 1)
attention_bwd_bh_kernel(const bf16* __restrict__ qkv, const long long* __restrict__ mask,
                        const bf16* __restrict__ dctx, const float* __restrict__ lse,
                        const float* __restrict__ delta, bf16* __restrict__ dqkv, int B, int S, int SP, int nbuf,
                        DropCfg drop) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    const uint32_t s0 = static_cast<uint32_t>(__cvta_generic_to_shared(smem));
    const int buf_bytes = 4 * SP * 128;
    float* s_lse = reinterpret_cast<float*>(smem + nbuf * buf_bytes);
    float* s_del = s_lse + SP;
    float* s_mb = s_del + SP;
    const int items = B * NH;
    auto issue = [&](int item, int buf) {
        const int b = item / NH, h = item % NH;
        const long long base = (long long)b * S;
        const uint32_t sb = s0 + buf * buf_bytes;
        load_tile_n(sb, qkv, QKV_LD, base, h * HD, SP, S, blockDim.x);
        load_tile_n(sb + SP * 128, qkv, QKV_LD, base, HID + h * HD, SP, S, blockDim.x);
        load_tile_n(sb + 2 * SP * 128, qkv, QKV_LD, base, 2 * HID + h * HD, SP, S, blockDim.x);
        load_tile_n(sb + 3 * SP * 128, dctx, HID, base, h * HD, SP, S, blockDim.x);
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    const int t = lane & 3, g = lane >> 2;
    int it = blockIdx.x, buf = 0;
    griddep_sync();
    if (it < items) issue(it, 0);
    for (; it < items; it += gridDim.x) {
        const int b = it / NH, h = it % NH;
        const long long base = (long long)b * S;
        const int nxt = it + gridDim.x;
        if (nbuf == 2 && nxt < items) {
            issue(nxt, buf ^ 1);
            asm volatile("cp.async.wait_group 1;" ::: "memory");
        } else {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        {
            const float* L = lse + ((long long)b * NH + h) * S;
            const float* Dl = delta + ((long long)b * NH + h) * S;
            for (int i = threadIdx.x; i < SP; i += blockDim.x) {
                s_lse[i] = i < S ? L[i] * LOG2E : INFINITY;     // +inf: padded rows contribute exp2(-inf) = 0
                s_del[i] = i < S ? Dl[i] : 0.f;
                s_mb[i] = i < S ? (mask[base + i] != 0 ? 0.f : MASK_LOG2) : -INFINITY;
            }
        }
        __syncthreads();
        const uint32_t sQ = s0 + buf * buf_bytes, sK = sQ + SP * 128, sV = sK + SP * 128, sdO = sV + SP * 128;
        const uint32_t hkey = drop_head_key(drop.key, b * NH + h);
        // ---------------- phase A: dQ, warp = 16 query rows, loop over keys
        for (int q0 = warp * 16; q0 < S; q0 += nw * 16) {
            uint32_t qf[4][4], dof[4][4];
            load_a_frags(sQ, q0, lane, qf);
            load_a_frags(sdO, q0, lane, dof);
            const float lse_lo = s_lse[q0 + g], lse_hi = s_lse[q0 + g + 8];
            const float d_lo = s_del[q0 + g], d_hi = s_del[q0 + g + 8];
            float dq[8][4];
#pragma unroll
            for (int n = 0; n < 8; ++n) dq[n][0] = dq[n][1] = dq[n][2] = dq[n][3] = 0.f;
            for (int kc = 0; kc < SP; kc += KC) {
                float sc[4][4], dp[4][4];
#pragma unroll
                for (int n = 0; n < 4; ++n) {
                    sc[n][0] = sc[n][1] = sc[n][2] = sc[n][3] = 0.f;
                    dp[n][0] = dp[n][1] = dp[n][2] = dp[n][3] = 0.f;
                }
                mma_nk_t<2>(sc, qf, sK, kc, lane);
                mma_nk_t<2>(dp, dof, sV, kc, lane);
#pragma unroll
                for (int n = 0; n < 4; ++n) {
                    const float2 bb = *reinterpret_cast<const float2*>(s_mb + kc + n * 8 + 2 * t);
                    const float p0 = ex2a(fmaf(sc[n][0], SCALE_LOG2, bb.x) - lse_lo);
                    const float p1 = ex2a(fmaf(sc[n][1], SCALE_LOG2, bb.y) - lse_lo);
                    const float p2 = ex2a(fmaf(sc[n][2], SCALE_LOG2, bb.x) - lse_hi);
                    const float p3 = ex2a(fmaf(sc[n][3], SCALE_LOG2, bb.y) - lse_hi);
                    if (drop.thresh) {        // dP = dropout'(dO V^T): same mask and scale as the forward pass
#pragma unroll
                        for (int e = 0; e < 4; e += 2) {
                            const uint32_t idx = (uint32_t)(q0 + g + (e >> 1) * 8) * (uint32_t)S + kc + n * 8 + 2 * t;
                            bool k0, k1;
                            drop_keep2(hkey, idx, drop.thresh, k0, k1);
                            dp[n][e] = k0 ? dp[n][e] * drop.scale : 0.f;
                            dp[n][e + 1] = k1 ? dp[n][e + 1] * drop.scale : 0.f;
                        }
                    }
                    sc[n][0] = p0 * (dp[n][0] - d_lo) * 0.125f; sc[n][1] = p1 * (dp[n][1] - d_lo) * 0.125f;
                    sc[n][2] = p2 * (dp[n][2] - d_hi) * 0.125f; sc[n][3] = p3 * (dp[n][3] - d_hi) * 0.125f;
                }
                mma_kn_t<2>(dq, sc, sK, kc, lane);
            }
            store_c_bf16(dqkv, QKV_LD, base + q0, h * HD, lane, dq, S, q0);
        }
        // ---------------- phase B: dK, dV, warp = 16 key rows, loop over queries
        for (int k0 = warp * 16; k0 < S; k0 += nw * 16) {
            uint32_t kf[4][4], vf[4][4];
            load_a_frags(sK, k0, lane, kf);
            load_a_frags(sV, k0, lane, vf);
            const float mb_lo = s_mb[k0 + g], mb_hi = s_mb[k0 + g + 8];
            float dk[8][4], dv[8][4];
#pragma unroll
            for (int n = 0; n < 8; ++n) {
                dk[n][0] = dk[n][1] = dk[n][2] = dk[n][3] = 0.f;
                dv[n][0] = dv[n][1] = dv[n][2] = dv[n][3] = 0.f;
            }
            for (int qc = 0; qc < SP; qc += KC) {
                float st[4][4], dpt[4][4];
#pragma unroll
                for (int n = 0; n < 4; ++n) {
                    st[n][0] = st[n][1] = st[n][2] = st[n][3] = 0.f;
                    dpt[n][0] = dpt[n][1] = dpt[n][2] = dpt[n][3] = 0.f;
                }
                mma_nk_t<2>(st, kf, sQ, qc, lane);        // S^T[key, q]
#pragma unroll
                for (int n = 0; n < 4; ++n) {
                    const float2 ll = *reinterpret_cast<const float2*>(s_lse + qc + n * 8 + 2 * t);
                    st[n][0] = ex2a(fmaf(st[n][0], SCALE_LOG2, mb_lo) - ll.x);
                    st[n][1] = ex2a(fmaf(st[n][1], SCALE_LOG2, mb_lo) - ll.y);
                    st[n][2] = ex2a(fmaf(st[n][2], SCALE_LOG2, mb_hi) - ll.x);
                    st[n][3] = ex2a(fmaf(st[n][3], SCALE_LOG2, mb_hi) - ll.y);
                }
                mma_nk_t<2>(dpt, vf, sdO, qc, lane);      // dP^T[key, q]
                uint32_t km = 0xFFFFu;                     // keep bits of this thread's 16 elements
                if (drop.thresh) {
                    km = 0;
#pragma unroll
                    for (int n = 0; n < 4; ++n)
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const uint32_t idx = (uint32_t)(qc + n * 8 + 2 * t + (e & 1)) * (uint32_t)S + k0 + g + (e >> 1) * 8;
                            km |= (drop_keep(hkey, idx, drop.thresh) ? 1u : 0u) << (n * 4 + e);
                        }
                }
                // dS^T = P^T (dropout'(dP^T) - delta) / 8 goes into dpt; then P^T is dropped in place for dV
#pragma unroll
                for (int n = 0; n < 4; ++n) {
                    const float2 ee = *reinterpret_cast<const float2*>(s_del + qc + n * 8 + 2 * t);
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const bool kp = (km >> (n * 4 + e)) & 1u;
                        const float dpe = drop.thresh ? (kp ? dpt[n][e] * drop.scale : 0.f) : dpt[n][e];
                        dpt[n][e] = st[n][e] * (dpe - ((e & 1) ? ee.y : ee.x)) * 0.125f;
                        if (drop.thresh) st[n][e] = kp ? st[n][e] * drop.scale : 0.f;
                    }
                }
                mma_kn_t<2>(dv, st, sdO, qc, lane);       // dV += dropout(P)^T dO
                mma_kn_t<2>(dk, dpt, sQ, qc, lane);       // dK += dS^T Q
            }
            store_c_bf16(dqkv, QKV_LD, base + k0, HID + h * HD, lane, dk, S, k0);
            store_c_bf16(dqkv, QKV_LD, base + k0, 2 * HID + h * HD, lane, dv, S, k0);
        }
        __syncthreads();          // everyone is done with this buffer and the float arrays
        if (nbuf == 1 && nxt < items) issue(nxt, 0);     // single buffer (S > 208): the reload is not overlapped
        buf ^= (nbuf == 2);
    }
}

// C[16 x 64(n)] += A[16 x 16 NK] * B with ready-made A fragments; B tile rows k0.. are the k index ([k][n])
template <int NK>
__device__ __forceinline__ void mma_kn_a(float c[8][4], const uint32_t (*a)[4], uint32_t tile, int k0, int lane) {
#pragma unroll
    for (int kk = 0; kk < NK; ++kk) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            uint32_t r[4];
            ldsm_x4_t(tile + tile_off(k0 + kk * 16 + (lane & 7) + (((lane >> 3) & 1) << 3), q * 2 + (lane >> 4)), r);
            mma16816(c[2 * q], a[kk], r[0], r[1]);
            mma16816(c[2 * q + 1], a[kk], r[2], r[3]);
        }
    }
}

// Backward for S <= 160 with P and dS cached: phase A (warp = 16 query rows) computes P = softmax probabilities and
// dS once, accumulates dQ, and parks dropout(P) and dS as bf16 [q][k] matrices in shared memory; phase B (warp = 16
// key rows) reads them back TRANSPOSED (ldmatrix.trans) as the A operands of dV += dropout(P)^T dO and
// dK += dS^T Q.  Five S x S x 64 products instead of the seven of the recompute kernel, and no exp / dropout hash in
// phase B.  Shared memory: Q K V dO tiles, the two [SP][SP] bf16 matrices (row pitch SP * 2 + 16 bytes: an odd number
// of 16-byte units keeps the transposed 8 x 8 loads bank-conflict free) and three float vectors -- a single buffer,
// so the next head's K, V are prefetched under phase B (they are dead by then) and its Q, dO after it.
__global__ void __launch_bounds__(BH_MAX_WARPS * 32, 1)
attention_bwd_bh_cached_kernel(const bf16* __restrict__ qkv, const long long* __restrict__ mask,
                               const bf16* __restrict__ dctx, const float* __restrict__ lse,
                               const bf16* __restrict__ ctx, bf16* __restrict__ dqkv, int B, int S, int SP,
                               DropCfg drop) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    const uint32_t sQ = static_cast<uint32_t>(__cvta_generic_to_shared(smem));
    const uint32_t sK = sQ + SP * 128, sV = sK + SP * 128, sdO = sV + SP * 128;
    const int PST = SP * 2 + 16;
    const uint32_t sP = sdO + SP * 128, sDS = sP + SP * PST, sO = sDS + SP * PST;
    float* s_lse = reinterpret_cast<float*>(smem + 5 * SP * 128 + 2 * SP * PST);
    float* s_del = s_lse + SP;
    float* s_mb = s_del + SP;
    const int items = B * NH;
    const int S16 = (S + 15) / 16 * 16;
    const int t = lane & 3, g = lane >> 2;
    auto issue_qdo = [&](int item) {
        const int b = item / NH, h = item % NH;
        load_tile_n(sQ, qkv, QKV_LD, (long long)b * S, h * HD, SP, S, blockDim.x);
        load_tile_n(sdO, dctx, HID, (long long)b * S, h * HD, SP, S, blockDim.x);
        load_tile_n(sO, ctx, HID, (long long)b * S, h * HD, SP, S, blockDim.x);
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    auto issue_kv = [&](int item) {
        const int b = item / NH, h = item % NH;
        load_tile_n(sK, qkv, QKV_LD, (long long)b * S, HID + h * HD, SP, S, blockDim.x);
        load_tile_n(sV, qkv, QKV_LD, (long long)b * S, 2 * HID + h * HD, SP, S, blockDim.x);
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    griddep_sync();
    int it = blockIdx.x;
    if (it < items) { issue_qdo(it); issue_kv(it); }
    for (; it < items; it += gridDim.x) {
        const int b = it / NH, h = it % NH;
        const long long base = (long long)b * S;
        const int nxt = it + gridDim.x;
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        {
            const float* L = lse + ((long long)b * NH + h) * S;
            for (int i = threadIdx.x; i < SP; i += blockDim.x) {
                s_lse[i] = i < S ? L[i] * LOG2E : INFINITY;     // +inf: padded rows contribute exp2(-inf) = 0
                s_mb[i] = i < S ? (mask[base + i] != 0 ? 0.f : MASK_LOG2) : -INFINITY;
            }
            // delta[q] = sum_d dO[q,d] * O[q,d].  Same (row, 16-byte chunk) -> thread mapping as load_tile_n, so
            // every thread reads back only what its own cp.async brought in (complete after the wait above); the
            // 8 chunks of a row sit in 8 consecutive lanes.  Rows >= S were zero-filled.
            for (int i = threadIdx.x; i < SP * 8; i += blockDim.x) {
                const int r = i >> 3, c = i & 7;
                float a[8], d[8];
                load8_bf16(reinterpret_cast<const bf16*>(smem + (sO - sQ) + tile_off(r, c)), a);
                load8_bf16(reinterpret_cast<const bf16*>(smem + (sdO - sQ) + tile_off(r, c)), d);
                float acc = 0.f;
#pragma unroll
                for (int k = 0; k < 8; ++k) acc += a[k] * d[k];
                acc += __shfl_xor_sync(0xffffffffu, acc, 1);
                acc += __shfl_xor_sync(0xffffffffu, acc, 2);
                acc += __shfl_xor_sync(0xffffffffu, acc, 4);
                if (c == 0) s_del[r] = acc;
            }
        }
        __syncthreads();
        const uint32_t hkey = drop_head_key(drop.key, b * NH + h);
        // ---------------- phase A: P, dS (parked in shared memory) and dQ; warp = 16 query rows
        for (int q0 = warp * 16; q0 < S; q0 += nw * 16) {
            uint32_t qf[4][4], dof[4][4];
            load_a_frags(sQ, q0, lane, qf);
            load_a_frags(sdO, q0, lane, dof);
            const float lse_lo = s_lse[q0 + g], lse_hi = s_lse[q0 + g + 8];
            const float d_lo = s_del[q0 + g], d_hi = s_del[q0 + g + 8];
            float dq[8][4];
#pragma unroll
            for (int n = 0; n < 8; ++n) dq[n][0] = dq[n][1] = dq[n][2] = dq[n][3] = 0.f;
            for (int kc = 0; kc < SP; kc += KC) {
                float sc[4][4], dp[4][4];
#pragma unroll
                for (int n = 0; n < 4; ++n) {
                    sc[n][0] = sc[n][1] = sc[n][2] = sc[n][3] = 0.f;
                    dp[n][0] = dp[n][1] = dp[n][2] = dp[n][3] = 0.f;
                }
                mma_nk_t<2>(sc, qf, sK, kc, lane);
                mma_nk_t<2>(dp, dof, sV, kc, lane);
#pragma unroll
                for (int n = 0; n < 4; ++n) {
                    const float2 bb = *reinterpret_cast<const float2*>(s_mb + kc + n * 8 + 2 * t);
                    float pr[4];
                    pr[0] = ex2a(fmaf(sc[n][0], SCALE_LOG2, bb.x) - lse_lo);
                    pr[1] = ex2a(fmaf(sc[n][1], SCALE_LOG2, bb.y) - lse_lo);
                    pr[2] = ex2a(fmaf(sc[n][2], SCALE_LOG2, bb.x) - lse_hi);
                    pr[3] = ex2a(fmaf(sc[n][3], SCALE_LOG2, bb.y) - lse_hi);
                    float pd[4] = {pr[0], pr[1], pr[2], pr[3]};          // dropout(P): what dV sees
                    if (drop.thresh) {
#pragma unroll
                        for (int e = 0; e < 4; e += 2) {
                            const uint32_t idx = (uint32_t)(q0 + g + (e >> 1) * 8) * (uint32_t)S + kc + n * 8 + 2 * t;
                            bool kp[2];
                            drop_keep2(hkey, idx, drop.thresh, kp[0], kp[1]);
#pragma unroll
                            for (int u = 0; u < 2; ++u) {
                                pd[e + u] = kp[u] ? pr[e + u] * drop.scale : 0.f;
                                dp[n][e + u] = kp[u] ? dp[n][e + u] * drop.scale : 0.f;
                            }
                        }
                    }
                    sc[n][0] = pr[0] * (dp[n][0] - d_lo) * 0.125f; sc[n][1] = pr[1] * (dp[n][1] - d_lo) * 0.125f;
                    sc[n][2] = pr[2] * (dp[n][2] - d_hi) * 0.125f; sc[n][3] = pr[3] * (dp[n][3] - d_hi) * 0.125f;
                    const uint32_t off_lo = (q0 + g) * PST + (kc + n * 8 + 2 * t) * 2, off_hi = off_lo + 8 * PST;
                    ptx::st_shared_b32(sP + off_lo, pack_bf16(pd[0], pd[1]));
                    ptx::st_shared_b32(sP + off_hi, pack_bf16(pd[2], pd[3]));
                    ptx::st_shared_b32(sDS + off_lo, pack_bf16(sc[n][0], sc[n][1]));
                    ptx::st_shared_b32(sDS + off_hi, pack_bf16(sc[n][2], sc[n][3]));
                }
                mma_kn_t<2>(dq, sc, sK, kc, lane);
            }
            store_c_bf16(dqkv, QKV_LD, base + q0, h * HD, lane, dq, S, q0);
        }
        __syncthreads();                      // P and dS are complete; this head's K and V tiles are dead
        if (nxt < items) issue_kv(nxt);
        // ---------------- phase B: dV = dropout(P)^T dO, dK = dS^T Q; warp = 16 key rows
        for (int k0 = warp * 16; k0 < S; k0 += nw * 16) {
            float dk[8][4], dv[8][4];
#pragma unroll
            for (int n = 0; n < 8; ++n) {
                dk[n][0] = dk[n][1] = dk[n][2] = dk[n][3] = 0.f;
                dv[n][0] = dv[n][1] = dv[n][2] = dv[n][3] = 0.f;
            }
            const uint32_t a_off = ((lane & 7) + ((lane >> 4) << 3)) * PST + (k0 + (((lane >> 3) & 1) << 3)) * 2;
            for (int qc = 0; qc < S16; qc += 16) {
                uint32_t ap[1][4], ad[1][4];
                ldsm_x4_t(sP + qc * PST + a_off, ap[0]);
                ldsm_x4_t(sDS + qc * PST + a_off, ad[0]);
                mma_kn_a<1>(dv, ap, sdO, qc, lane);
                mma_kn_a<1>(dk, ad, sQ, qc, lane);
            }
            store_c_bf16(dqkv, QKV_LD, base + k0, HID + h * HD, lane, dk, S, k0);
            store_c_bf16(dqkv, QKV_LD, base + k0, 2 * HID + h * HD, lane, dv, S, k0);
        }
        __syncthreads();                      // Q, dO, P, dS are dead
        if (nxt < items) issue_qdo(nxt);
    }
}

// warps per CTA for the per-head kernels: every warp gets the same number of 16-row tiles
int bh_warps(int S) {
    const int tiles = (S + 15) / 16;
    const int rounds = (tiles + BH_MAX_WARPS - 1) / BH_MAX_WARPS;
    return (tiles + rounds - 1) / rounds;
}

int attn_check(int B, int S, int* S_pad) {
    UC2_REQUIRE(B > 0 && S > 0, UC2_ERR_ARG, "attention: bad shape B=%d S=%d", B, S);
    UC2_REQUIRE(S <= 512, UC2_ERR_UNSUPPORTED, "attention: S=%d > 512 (max_position_embeddings cap)", S);
    *S_pad = (S + TILE - 1) / TILE * TILE;
    return UC2_OK;
}

template <typename K>
int set_smem(K kern, int bytes) {
    UC2_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    return UC2_OK;
}

}  // namespace
}  // namespace uc2

using namespace uc2;

extern "C" UC2_API int uc2_attention_fwd(const void* qkv, const long long* attn_mask, void* ctx, float* lse, int B,
                                         int S, void* stream) {
    return uc2_attention_fwd_dropout(qkv, attn_mask, ctx, lse, B, S, 0u, 0u, 1.f, stream);
}

extern "C" UC2_API int uc2_attention_fwd_dropout(const void* qkv, const long long* attn_mask, void* ctx, float* lse,
                                                 int B, int S, unsigned int drop_key, unsigned int drop_thresh,
                                                 float drop_scale, void* stream) {
    if (int rc = require_sm100()) return rc;
    UC2_REQUIRE(drop_thresh < 65536u, UC2_ERR_ARG, "attention_fwd: drop_thresh must be < 65536");
    UC2_REQUIRE(drop_thresh == 0 || S <= 256, UC2_ERR_UNSUPPORTED, "attention dropout is implemented for S <= 256 (S=%d)", S);
    const DropCfg drop = {drop_key, drop_thresh, drop_scale};
    UC2_REQUIRE(qkv && attn_mask && ctx && lse, UC2_ERR_ARG, "attention_fwd: null pointer");
    UC2_REQUIRE(aligned16(qkv) && aligned16(ctx), UC2_ERR_ARG, "attention_fwd: qkv/ctx must be 16-byte aligned");
    int S_pad;
    if (int rc = attn_check(B, S, &S_pad)) return rc;
    // tcgen05 / TMEM kernels (attention_tc.cu) unless switched off.  With attention dropout the forward and the backward
    // have to regenerate the same mask, and the two kernel families draw it from different counter streams: the
    // forward only goes there when the backward of this shape will too (attn_tc_bwd_serves).
    // (The choice must not depend on pointer alignment, or the two passes could disagree: the tcgen05 entry points
    // reject a ctx / dqkv that is not 32-byte aligned.)
    if (attn_tc_enabled() && S <= 256 && (drop_thresh == 0 || attn_tc_bwd_serves(S)))
        return uc2_attention_fwd_tc(qkv, attn_mask, ctx, lse, B, S, drop_key, drop_thresh, drop_scale, stream);
    ProfScope prof((cudaStream_t)stream, 1, 4.0 * B * NH * (double)S * S * HD);
    if (S <= 256) {
        const int SP = (S + 31) / 32 * 32;
        const int smem_bh = 3 * SP * 128 + SP * 4;
        if (int rc = set_smem(attention_fwd_bh_kernel, smem_bh)) return rc;
        launch_pdl(attention_fwd_bh_kernel, dim3(NH, B), dim3(bh_warps(S) * 32), smem_bh, (cudaStream_t)stream, 1,
                   (const bf16*)qkv, attn_mask, (bf16*)ctx, lse, S, SP, drop);
        return check_last("attention_fwd_bh_kernel");
    }
    const int smem = TILE * 128 + 2 * S_pad * 128 + S_pad * 4;
    if (int rc = set_smem(attention_fwd_kernel, smem)) return rc;
    attention_fwd_kernel<<<dim3(S_pad / TILE, NH, B), ATT_THREADS, smem, (cudaStream_t)stream>>>(
        (const bf16*)qkv, attn_mask, (bf16*)ctx, lse, S, S_pad);
    return check_last("attention_fwd_kernel");
}

extern "C" UC2_API int uc2_attention_bwd(const void* qkv, const long long* attn_mask, const void* ctx,
                                         const void* dctx, const float* lse, float* delta_ws, void* dqkv, int B,
                                         int S, void* stream) {
    return uc2_attention_bwd_dropout(qkv, attn_mask, ctx, dctx, lse, delta_ws, dqkv, B, S, 0u, 0u, 1.f, stream);
}

extern "C" UC2_API int uc2_attention_bwd_dropout(const void* qkv, const long long* attn_mask, const void* ctx,
                                                 const void* dctx, const float* lse, float* delta_ws, void* dqkv,
                                                 int B, int S, unsigned int drop_key, unsigned int drop_thresh,
                                                 float drop_scale, void* stream) {
    if (int rc = require_sm100()) return rc;
    UC2_REQUIRE(drop_thresh < 65536u, UC2_ERR_ARG, "attention_bwd: drop_thresh must be < 65536");
    UC2_REQUIRE(drop_thresh == 0 || S <= 256, UC2_ERR_UNSUPPORTED, "attention dropout is implemented for S <= 256 (S=%d)", S);
    const DropCfg drop = {drop_key, drop_thresh, drop_scale};
    UC2_REQUIRE(qkv && attn_mask && ctx && dctx && lse && delta_ws && dqkv, UC2_ERR_ARG, "attention_bwd: null pointer");
    UC2_REQUIRE(aligned16(qkv) && aligned16(ctx) && aligned16(dctx) && aligned16(dqkv), UC2_ERR_ARG,
                "attention_bwd: tensors must be 16-byte aligned");
    int S_pad;
    if (int rc = attn_check(B, S, &S_pad)) return rc;
    if (attn_tc_enabled() && attn_tc_bwd_serves(S))
        return uc2_attention_bwd_tc(qkv, attn_mask, ctx, dctx, lse, dqkv, B, S, drop_key, drop_thresh, drop_scale, stream);
    cudaStream_t s = (cudaStream_t)stream;
    const long long rows = (long long)B * S;
    ProfScope prof(s, 1, 10.0 * B * NH * (double)S * S * HD);      // 5 S x S x 64 products (7 computed)
    auto launch_delta = [&]() {
        launch_pdl(attention_delta_kernel, dim3((unsigned)((rows + 7) / 8)), dim3(256), 0, s, 1, (const bf16*)ctx,
                   (const bf16*)dctx, delta_ws, B, S);
        return check_last("attention_delta_kernel");
    };
    if (S <= 256) {
        const int SP = (S + 31) / 32 * 32;
        // Q K V dO O tiles + P, dS + three float vectors; delta is computed in the kernel from the O and dO tiles
        const int smem_cached = 5 * SP * 128 + 2 * SP * (SP * 2 + 16) + 3 * SP * 4;
        if (smem_cached <= 225 * 1024) {
            if (int rc = set_smem(attention_bwd_bh_cached_kernel, smem_cached)) return rc;
            const int items = B * NH;
            const int grid = items < num_sms() ? items : num_sms();
            launch_pdl(attention_bwd_bh_cached_kernel, dim3(grid), dim3(bh_warps(S) * 32), smem_cached, s, 1,
                       (const bf16*)qkv, attn_mask, (const bf16*)dctx, lse, (const bf16*)ctx, (bf16*)dqkv, B, S, SP,
                       drop);
            return check_last("attention_bwd_bh_cached_kernel");
        }
        if (int rc = launch_delta()) return rc;
        const int nbuf = (2 * 4 * SP * 128 + 3 * SP * 4 <= 220 * 1024) ? 2 : 1;
        const int smem_bh = nbuf * 4 * SP * 128 + 3 * SP * 4;
        if (int rc = set_smem(attention_bwd_bh_kernel, smem_bh)) return rc;
        const int items = B * NH;
        const int grid = items < num_sms() ? items : num_sms();
        launch_pdl(attention_bwd_bh_kernel, dim3(grid), dim3(bh_warps(S) * 32), smem_bh, s, 1, (const bf16*)qkv, attn_mask,
                   (const bf16*)dctx, lse, delta_ws, (bf16*)dqkv, B, S, SP, nbuf, drop);
        return check_last("attention_bwd_bh_kernel");
    }
    if (int rc = launch_delta()) return rc;
    const int smem_dq = 2 * TILE * 128 + 2 * S_pad * 128 + S_pad * 4;
    const int smem_dkv = 2 * TILE * 128 + 2 * S_pad * 128 + 2 * S_pad * 4;
    if (int rc = set_smem(attention_bwd_dq_kernel, smem_dq)) return rc;
    if (int rc = set_smem(attention_bwd_dkv_kernel, smem_dkv)) return rc;
    attention_bwd_dq_kernel<<<dim3(S_pad / TILE, NH, B), ATT_THREADS, smem_dq, s>>>(
        (const bf16*)qkv, attn_mask, (const bf16*)dctx, lse, delta_ws, (bf16*)dqkv, S, S_pad);
    if (int rc = check_last("attention_bwd_dq_kernel")) return rc;
    attention_bwd_dkv_kernel<<<dim3(S_pad / TILE, NH, B), ATT_THREADS, smem_dkv, s>>>(
        (const bf16*)qkv, attn_mask, (const bf16*)dctx, lse, delta_ws, (bf16*)dqkv, S, S_pad);
    return check_last("attention_bwd_dkv_kernel");
}
