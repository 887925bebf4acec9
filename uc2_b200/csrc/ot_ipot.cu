// WRA optimal-transport distance: cosine cost + 50 IPOT iterations + trace(C.T), one CTA per sample with
// the whole working set (C, A, T: 3 x N x M fp32) resident in shared memory, plus its backward.
// Replaces model/ot.py:8-82 (cost_matrix_cosine, ipot, trace, optimal_transport_dist) and the scatter
// un-pack of forward_itm (model/model.py:703-716), which is folded into the row addressing here.
// The reference launches ~350 small kernels per call (50 x {mul, 2 bmm, 2 reciprocals, ...}); this is one
// launch, latency-bound by the 50 sequential iterations (neither HBM nor tensor roofline applies).
#include "common.cuh"

namespace uc2 {
namespace {

constexpr int OT_THREADS = 512;
constexpr int KC = 64;            // hidden columns staged per step of the cosine matmul
constexpr int KLD = KC + 1;

struct OtParams {
    const bf16* seq;              // [B,S,768] packed encoder output
    const long long* scatter;     // [B,S] ot_scatter: packed row j -> context position
    const unsigned char* txt_pad; // [B,M]
    const unsigned char* img_pad; // [B,N]
    int B, S, M, N, tl;           // tl = input_ids.size(1): image context rows start here
    float beta; int iters; int k;
    float* C_out; float* T_out;   // [B,M,N] and [B,N,M] saved for backward (may be null)
    float* dist;                  // [B]
};

__device__ __forceinline__ float block_reduce_sum(float v, float* red) {
    v = warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    float r = 0.f;
    for (int w = 0; w < OT_THREADS / 32; ++w) r += red[w];
    return r;
}

// inv[c] = packed row j with scatter[b,j] == c (or -1), for c < tl + N
__device__ __forceinline__ void build_inverse(const OtParams& p, int b, int* inv, int n_ctx) {
    for (int c = threadIdx.x; c < n_ctx; c += OT_THREADS) inv[c] = -1;
    __syncthreads();
    for (int j = threadIdx.x; j < p.S; j += OT_THREADS) {
        const long long c = p.scatter[(long long)b * p.S + j];
        if (c >= 0 && c < n_ctx) inv[c] = j;
    }
    __syncthreads();
}

// 1 / max(||row||, eps) for the M text rows then the N image rows (F.normalize, ot.py:14-15)
__device__ __forceinline__ void row_inv_norms(const OtParams& p, int b, const int* inv, float* inorm) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int r = warp; r < p.M + p.N; r += OT_THREADS / 32) {
        const int c = r < p.M ? r : p.tl + (r - p.M);
        const int j = inv[c];
        float s = 0.f;
        if (j >= 0) {
            const bf16* row = p.seq + ((long long)b * p.S + j) * HID;
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                float v[8];
                load8_bf16(row + i * 256 + lane * 8, v);
#pragma unroll
                for (int k = 0; k < 8; ++k) s += v[k] * v[k];
            }
        }
        s = warp_sum(s);
        if (lane == 0) inorm[r] = 1.f / fmaxf(sqrtf(s), 1e-5f);
    }
    __syncthreads();
}

// stage columns [k0, k0+KC) of the normalised rows into shared memory as fp32
__device__ __forceinline__ void stage_chunk(const OtParams& p, int b, const int* inv, const float* inorm, int k0,
                                            float* xs, float* ys) {
    const int rows = p.M + p.N;
    for (int i = threadIdx.x; i < rows * (KC / 8); i += OT_THREADS) {
        const int r = i / (KC / 8), c8 = (i % (KC / 8)) * 8;
        const int c = r < p.M ? r : p.tl + (r - p.M);
        const int j = inv[c];
        float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (j >= 0) load8_bf16(p.seq + ((long long)b * p.S + j) * HID + k0 + c8, v);
        float* dst = (r < p.M ? xs + r * KLD : ys + (r - p.M) * KLD) + c8;
        const float s = inorm[r];
#pragma unroll
        for (int k = 0; k < 8; ++k) dst[k] = v[k] * s;
    }
}

__global__ void __launch_bounds__(OT_THREADS)
ot_ipot_fwd_kernel(const OtParams p) {
    extern __shared__ __align__(16) float sm[];
    const int b = blockIdx.x;
    const int M = p.M, N = p.N, MN = M * N;
    // layout: Cs[N*M] | region1 (A[N*M], T[N*M]  or  xs, ys chunk buffers) | vectors | inv
    float* Cs = sm;                                  // stored [n][m] like T
    float* r1 = Cs + MN;
    const int r1_floats = max(2 * MN, (M + N) * KLD);
    float* As = r1;
    float* Ts = r1 + MN;
    float* xs = r1;
    float* ys = r1 + M * KLD;
    float* inorm = r1 + r1_floats;                   // [M+N]
    float* sigma = inorm + M + N;                    // [M]
    float* delta = sigma + M;                        // [N]
    float* xmask = delta + N;                        // [M]
    float* ymask = xmask + M;                        // [N]
    float* red = ymask + N;                          // [32]
    int* inv = reinterpret_cast<int*>(red + 32);     // [tl + N]
    const unsigned char* tp = p.txt_pad + (long long)b * M;
    const unsigned char* ip = p.img_pad + (long long)b * N;

    build_inverse(p, b, inv, p.tl + N);
    row_inv_norms(p, b, inv, inorm);

    // ---- cosine similarity: 4x4 register tiles over (m, n), K staged in chunks ----
    const int tm_n = (M + 3) / 4, tn_n = (N + 3) / 4;
    const int tiles = tm_n * tn_n;
    // each thread owns up to 2 tiles (M,N <= 128 -> <= 1024 tiles)
    float acc[2][16];
#pragma unroll
    for (int t = 0; t < 2; ++t)
#pragma unroll
        for (int i = 0; i < 16; ++i) acc[t][i] = 0.f;
    for (int k0 = 0; k0 < HID; k0 += KC) {
        __syncthreads();
        stage_chunk(p, b, inv, inorm, k0, xs, ys);
        __syncthreads();
#pragma unroll
        for (int t = 0; t < 2; ++t) {
            const int tile = threadIdx.x + t * OT_THREADS;
            if (tile >= tiles) break;
            const int m0 = (tile % tm_n) * 4, n0 = (tile / tm_n) * 4;
            for (int k = 0; k < KC; ++k) {
                float xv[4], yv[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    xv[i] = m0 + i < M ? xs[(m0 + i) * KLD + k] : 0.f;
                    yv[i] = n0 + i < N ? ys[(n0 + i) * KLD + k] : 0.f;
                }
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[t][i * 4 + j] += xv[i] * yv[j];
            }
        }
    }
    __syncthreads();
    // cost = 1 - cos, zero on the joint pad (ot.py:17, 71-72)
#pragma unroll
    for (int t = 0; t < 2; ++t) {
        const int tile = threadIdx.x + t * OT_THREADS;
        if (tile >= tiles) break;
        const int m0 = (tile % tm_n) * 4, n0 = (tile / tm_n) * 4;
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int m = m0 + i, n = n0 + j;
                if (m < M && n < N) {
                    const bool pad = tp[m] || ip[n];
                    Cs[n * M + m] = pad ? 0.f : 1.f - acc[t][i * 4 + j];
                }
            }
    }
    __syncthreads();     // xs/ys are dead from here: region1 becomes A and T

    // ---- IPOT (ot.py:32-63) ----
    int npx = 0, npy = 0;
    for (int m = 0; m < M; ++m) npx += tp[m] ? 1 : 0;
    for (int n = 0; n < N; ++n) npy += ip[n] ? 1 : 0;
    const float x_len = (float)(M - npx), y_len = (float)(N - npy);
    for (int i = threadIdx.x; i < MN; i += OT_THREADS) {
        const int n = i / M, m = i % M;
        const bool pad = tp[m] || ip[n];
        As[i] = pad ? 0.f : __expf(-Cs[i] / p.beta);
        Ts[i] = pad ? 0.f : 1.f;
    }
    for (int m = threadIdx.x; m < M; m += OT_THREADS) {
        sigma[m] = tp[m] ? 0.f : 1.f / x_len;
        xmask[m] = tp[m] ? 1e4f : 0.f;
    }
    for (int n = threadIdx.x; n < N; n += OT_THREADS) ymask[n] = ip[n] ? 1e4f : 0.f;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int NW = OT_THREADS / 32;
    for (int it = 0; it < p.iters; ++it) {
        // Q = A * T (kept in T's storage)
        for (int i = threadIdx.x; i < MN; i += OT_THREADS) Ts[i] = As[i] * Ts[i];
        __syncthreads();
        for (int kk = 0; kk < p.k; ++kk) {
            // delta[n] = 1 / (y_len * sum_m Q[n,m] sigma[m] + y_mask[n])
            for (int n = warp; n < N; n += NW) {
                float s = 0.f;
                for (int m = lane; m < M; m += 32) s += Ts[n * M + m] * sigma[m];
                s = warp_sum(s);
                if (lane == 0) delta[n] = 1.f / (y_len * s + ymask[n]);
            }
            __syncthreads();
            // sigma[m] = 1 / (x_len * sum_n delta[n] Q[n,m] + x_mask[m]); columns: consecutive threads -> consecutive m
            {
                constexpr int groups = OT_THREADS / 128;      // 4 n-slices x 128 columns (M <= 128)
                const int m = threadIdx.x & 127, gsl = threadIdx.x >> 7;
                float s = 0.f;
                if (m < M)
                    for (int n = gsl; n < N; n += groups) s += delta[n] * Ts[n * M + m];
                if (gsl == 0 && m < M) sigma[m] = 0.f;        // sigma is not read in this step
                __syncthreads();
                if (m < M) atomicAdd(sigma + m, s);           // 4-way cross-slice reduction
                __syncthreads();
                if (threadIdx.x < M) sigma[threadIdx.x] = 1.f / (x_len * sigma[threadIdx.x] + xmask[threadIdx.x]);
                __syncthreads();
            }
        }
        // T = delta[n] * Q * sigma[m]
        for (int i = threadIdx.x; i < MN; i += OT_THREADS) Ts[i] = delta[i / M] * Ts[i] * sigma[i % M];
        __syncthreads();
    }
    // distance = trace(C @ T) = sum_{m,n} C[m,n] T[n,m]; T is already zero on the joint pad
    float d = 0.f;
    for (int i = threadIdx.x; i < MN; i += OT_THREADS) d += Cs[i] * Ts[i];
    d = block_reduce_sum(d, red);
    if (threadIdx.x == 0) p.dist[b] = d;
    if (p.T_out)
        for (int i = threadIdx.x; i < MN; i += OT_THREADS) p.T_out[(long long)b * MN + i] = Ts[i];
    if (p.C_out)
        for (int i = threadIdx.x; i < MN; i += OT_THREADS) p.C_out[(long long)b * MN + i] = Cs[i];   // [n][m] layout
}

// backward: d(dist)/dC = T^T (T detached) -> through 1 - x^.y^ -> through F.normalize -> scatter to packed rows
__global__ void __launch_bounds__(OT_THREADS)
ot_ipot_bwd_kernel(const OtParams p, const float* __restrict__ ddist, bf16* __restrict__ dseq) {
    extern __shared__ __align__(16) float sm[];
    const int b = blockIdx.x;
    const int M = p.M, N = p.N, MN = M * N;
    float* Ts = sm;                                  // [n][m], pre-multiplied by -ddist[b], zero on pads
    float* xs = Ts + MN;                             // [M][KLD] normalised chunk
    float* ys = xs + M * KLD;                        // [N][KLD]
    float* inorm = ys + N * KLD;                     // [M+N]
    float* proj = inorm + M + N;                     // [M+N]: s_m = x^_m . dx^_m ; r_n likewise
    int* inv = reinterpret_cast<int*>(proj + M + N);
    const unsigned char* tp = p.txt_pad + (long long)b * M;
    const unsigned char* ip = p.img_pad + (long long)b * N;
    const float g = ddist[b];
    build_inverse(p, b, inv, p.tl + N);
    row_inv_norms(p, b, inv, inorm);
    for (int i = threadIdx.x; i < MN; i += OT_THREADS) {
        const int n = i / M, m = i % M;
        const bool pad = tp[m] || ip[n];
        Ts[i] = pad ? 0.f : -g * p.T_out[(long long)b * MN + i];
    }
    __syncthreads();
    // projections: x^_m . dx^_m = sum_n G[m,n] cos[m,n] with G = -g T^T, cos = 1 - C (non-pad entries)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int NW = OT_THREADS / 32;
    const float* Cg = p.C_out + (long long)b * MN;
    for (int r = warp; r < M + N; r += NW) {
        float s = 0.f;
        if (r < M) {
            for (int n = lane; n < N; n += 32) s += Ts[n * M + r] * (1.f - Cg[n * M + r]);
        } else {
            const int n = r - M;
            for (int m = lane; m < M; m += 32) s += Ts[n * M + m] * (1.f - Cg[n * M + m]);
        }
        s = warp_sum(s);
        if (lane == 0) proj[r] = s;
    }
    for (int k0 = 0; k0 < HID; k0 += KC) {
        __syncthreads();
        stage_chunk(p, b, inv, inorm, k0, xs, ys);
        __syncthreads();
        // dx[m,k] = (sum_n G[m,n] y^[n,k] - x^[m,k] s_m) / ||x_m||  ; dy[n,k] = (sum_m G[m,n] x^[m,k] - y^[n,k] r_n) / ||y_n||
        for (int i = threadIdx.x; i < (M + N) * (KC / 4); i += OT_THREADS) {
            const int r = i / (KC / 4), c4 = (i % (KC / 4)) * 4;
            const int ctx = r < M ? r : p.tl + (r - M);
            const int j = inv[ctx];
            if (j < 0) continue;
            float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
            const float* own;
            if (r < M) {
                for (int n = 0; n < N; ++n) {
                    const float t = Ts[n * M + r];
                    const float* y = ys + n * KLD + c4;
                    a0 += t * y[0]; a1 += t * y[1]; a2 += t * y[2]; a3 += t * y[3];
                }
                own = xs + r * KLD + c4;
            } else {
                const int n = r - M;
                for (int m = 0; m < M; ++m) {
                    const float t = Ts[n * M + m];
                    const float* x = xs + m * KLD + c4;
                    a0 += t * x[0]; a1 += t * x[1]; a2 += t * x[2]; a3 += t * x[3];
                }
                own = ys + n * KLD + c4;
            }
            const float s = proj[r], inn = inorm[r];
            // rows whose norm was clamped by eps get the plain 1/eps scaling (F.normalize backward); the
            // projection term is exact there as well because x^ = x/eps has |x^| < 1 -- negligible in practice.
            const float o0 = (a0 - own[0] * s) * inn, o1 = (a1 - own[1] * s) * inn;
            const float o2 = (a2 - own[2] * s) * inn, o3 = (a3 - own[3] * s) * inn;
            bf16* dst = dseq + ((long long)b * p.S + j) * HID + k0 + c4;
            *reinterpret_cast<uint2*>(dst) = make_uint2(pack_bf16(o0, o1), pack_bf16(o2, o3));
        }
    }
}

int ot_smem_fwd(int M, int N, int tl) {
    const int MN = M * N;
    const int r1 = 2 * MN > (M + N) * KLD ? 2 * MN : (M + N) * KLD;
    return (MN + r1 + (M + N) + M + N + M + N + 32) * 4 + (tl + N) * 4 + 64;
}
int ot_smem_bwd(int M, int N, int tl) {
    return (M * N + (M + N) * KLD + 2 * (M + N)) * 4 + (tl + N) * 4 + 64;
}

}  // namespace
}  // namespace uc2

using namespace uc2;

static int ot_fill(OtParams& p, const void* seq, const long long* scatter, const unsigned char* txt_pad,
                   const unsigned char* img_pad, int B, int S, int M, int N, int tl, float beta, int iters, int k) {
    UC2_REQUIRE(seq && scatter && txt_pad && img_pad && B > 0 && S > 0 && M > 0 && N > 0, UC2_ERR_ARG, "ot: bad args");
    UC2_REQUIRE(M <= tl, UC2_ERR_ARG, "ot: txt_pad is wider (%d) than the text block (%d)", M, tl);
    UC2_REQUIRE(M <= 128 && N <= 128, UC2_ERR_UNSUPPORTED, "ot: at most 128 tokens x 128 regions (got %d x %d)", M, N);
    p.seq = (const bf16*)seq; p.scatter = scatter; p.txt_pad = txt_pad; p.img_pad = img_pad;
    p.B = B; p.S = S; p.M = M; p.N = N; p.tl = tl; p.beta = beta; p.iters = iters; p.k = k;
    return UC2_OK;
}

extern "C" UC2_API int uc2_ot_ipot_fwd(const void* seq, const long long* ot_scatter, const unsigned char* txt_pad,
                                       const unsigned char* img_pad, int B, int S, int M, int N, int tl, float beta,
                                       int iterations, int k, float* dist, float* C_save, float* T_save, void* stream) {
    if (int rc = require_sm100()) return rc;
    OtParams p;
    if (int rc = ot_fill(p, seq, ot_scatter, txt_pad, img_pad, B, S, M, N, tl, beta, iterations, k)) return rc;
    UC2_REQUIRE(dist, UC2_ERR_ARG, "ot_ipot_fwd: dist is null");
    p.dist = dist; p.C_out = C_save; p.T_out = T_save;
    const int smem = ot_smem_fwd(M, N, tl);
    UC2_REQUIRE(smem <= 227 * 1024, UC2_ERR_UNSUPPORTED, "ot_ipot_fwd: working set %d B exceeds shared memory", smem);
    UC2_CUDA(cudaFuncSetAttribute(ot_ipot_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    ot_ipot_fwd_kernel<<<B, OT_THREADS, smem, (cudaStream_t)stream>>>(p);
    return check_last("ot_ipot_fwd_kernel");
}

extern "C" UC2_API int uc2_ot_ipot_bwd(const void* seq, const long long* ot_scatter, const unsigned char* txt_pad,
                                       const unsigned char* img_pad, int B, int S, int M, int N, int tl,
                                       const float* C_save, const float* T_save, const float* ddist, void* dseq,
                                       void* stream) {
    if (int rc = require_sm100()) return rc;
    OtParams p;
    if (int rc = ot_fill(p, seq, ot_scatter, txt_pad, img_pad, B, S, M, N, tl, 0.5f, 0, 1)) return rc;
    UC2_REQUIRE(C_save && T_save && ddist && dseq, UC2_ERR_ARG, "ot_ipot_bwd: null pointer");
    p.C_out = const_cast<float*>(C_save); p.T_out = const_cast<float*>(T_save); p.dist = nullptr;
    const int smem = ot_smem_bwd(M, N, tl);
    UC2_REQUIRE(smem <= 227 * 1024, UC2_ERR_UNSUPPORTED, "ot_ipot_bwd: working set %d B exceeds shared memory", smem);
    UC2_CUDA(cudaFuncSetAttribute(ot_ipot_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    ot_ipot_bwd_kernel<<<B, OT_THREADS, smem, (cudaStream_t)stream>>>(p, ddist, (bf16*)dseq);
    return check_last("ot_ipot_bwd_kernel");
}
