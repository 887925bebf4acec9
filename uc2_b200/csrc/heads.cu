// Small warp-level head kernels: narrow linear layers (ITM / rank outputs), pooler tanh backward,
// triplet ranking loss, small-class cross entropy, order-exact masked-row compaction and its scatter
// backward, fused log-softmax losses (KL / CE / MSE).
// Reference: model/layer.py:179-185 (pooler), model/itm.py:43-53 (rank loss), model/model.py:653-657
// (_compute_masked_hidden), 592-596 / 683-686 / 761-773 (losses), 698 + 732 (ITM head).
#include "common.cuh"

namespace uc2 {
namespace {

// ------------------------------------------------------------------------------------------------
// narrow linear: out[m,n] = sum_k x[m,k] W[n,k] + b[n],  N <= 8, K % 128 == 0;  warp per row
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
narrow_linear_fwd_kernel(const float* __restrict__ x, long long ldx, const float* __restrict__ W,
                         const float* __restrict__ b, float* __restrict__ out, int M, int N, int K) {
    const int lane = threadIdx.x & 31;
    const int m = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (m >= M) return;
    for (int n = 0; n < N; ++n) {
        float acc = 0.f;
        for (int k = lane * 4; k < K; k += 128) {
            const float4 a = *reinterpret_cast<const float4*>(x + m * ldx + k);
            const float4 w = *reinterpret_cast<const float4*>(W + (long long)n * K + k);
            acc += a.x * w.x + a.y * w.y + a.z * w.z + a.w * w.w;
        }
        acc = warp_sum(acc);
        if (lane == 0) out[(long long)m * N + n] = acc + (b ? b[n] : 0.f);
    }
}

// dx[m,k] = sum_n dy[m,n] W[n,k];   dW[n,k] += sum_m dy[m,n] x[m,k];   db[n] += sum_m dy[m,n]
__global__ void __launch_bounds__(256)
narrow_linear_bwd_kernel(const float* __restrict__ x, long long ldx, const float* __restrict__ W,
                         const float* __restrict__ dy, float* __restrict__ dx, long long lddx,
                         float* __restrict__ dW, float* __restrict__ db, int M, int N, int K) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= K) return;
    float w[8], gw[8];
    for (int n = 0; n < N; ++n) { w[n] = W[(long long)n * K + k]; gw[n] = 0.f; }
    for (int m = 0; m < M; ++m) {
        const float xv = x[m * ldx + k];
        float d = 0.f;
        for (int n = 0; n < N; ++n) {
            const float g = dy[(long long)m * N + n];
            d += g * w[n];
            gw[n] += g * xv;
        }
        if (dx) dx[m * lddx + k] = d;
    }
    for (int n = 0; n < N; ++n) atomicAdd(dW + (long long)n * K + k, gw[n]);
    if (k < N && db) {
        float s = 0.f;
        for (int m = 0; m < M; ++m) s += dy[(long long)m * N + k];
        atomicAdd(db + k, s);
    }
}

// dpre = dy * (1 - y^2) as bf16 (pooler tanh backward; y = tanh output fp32)
__global__ void tanh_bwd_kernel(const float* __restrict__ y, const float* __restrict__ dy, bf16* __restrict__ out,
                                long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = __float2bfloat16(dy[i] * (1.f - y[i] * y[i]));
}

// ------------------------------------------------------------------------------------------------
// triplet ranking loss (model/itm.py:43-53): s = sigmoid(score).view(-1, ss); loss = clamp(margin + neg - pos, 0)
// ------------------------------------------------------------------------------------------------
__global__ void rank_loss_fwd_kernel(const float* __restrict__ scores, float* __restrict__ loss, int groups, int ss,
                                     float margin) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= groups * (ss - 1)) return;
    const int g = i / (ss - 1), j = i % (ss - 1) + 1;
    const float pos = 1.f / (1.f + __expf(-scores[g * ss]));
    const float neg = 1.f / (1.f + __expf(-scores[g * ss + j]));
    loss[i] = fmaxf(margin + neg - pos, 0.f);
}
__global__ void rank_loss_bwd_kernel(const float* __restrict__ scores, const float* __restrict__ dloss,
                                     float* __restrict__ dscores, int groups, int ss, float margin) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= groups) return;
    const float pos = 1.f / (1.f + __expf(-scores[g * ss]));
    float dpos = 0.f;
    for (int j = 1; j < ss; ++j) {
        const float neg = 1.f / (1.f + __expf(-scores[g * ss + j]));
        const float d = (margin + neg - pos > 0.f) ? dloss[g * (ss - 1) + j - 1] : 0.f;
        dscores[g * ss + j] = d * neg * (1.f - neg);
        dpos -= d;
    }
    dscores[g * ss] = dpos * pos * (1.f - pos);
}

// ------------------------------------------------------------------------------------------------
// row-wise log-softmax losses over fp32 logits [n, C] (one CTA per row):
//   kind 0: cross entropy with int64 targets (ignore_index < 0: none)   F.cross_entropy(reduction='none')
//   kind 1: KL divergence vs soft targets [n, C], elementwise out [n, C]  F.kl_div(log_softmax, t, 'none')
// backward writes dlogits (fp32 [n,C]) given dloss.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float block_max(float v, float* red) {
    v = warp_max(v);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    float r = red[0];
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) r = fmaxf(r, red[w]);
    __syncthreads();
    return r;
}
__device__ __forceinline__ float block_sum(float v, float* red) {
    v = warp_sum(v);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    float r = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) r += red[w];
    __syncthreads();
    return r;
}

__global__ void __launch_bounds__(256)
softmax_loss_kernel(const float* __restrict__ logits, long long ld, int C, int kind,
                    const long long* __restrict__ targets, long long ignore_index,
                    const float* __restrict__ soft, float* __restrict__ loss, const float* __restrict__ dloss,
                    float* __restrict__ dlogits, float* __restrict__ lse_out) {
    __shared__ float red[8];
    const long long r = blockIdx.x;
    const float* x = logits + r * ld;
    float mx = -INFINITY;
    for (int c = threadIdx.x; c < C; c += blockDim.x) mx = fmaxf(mx, x[c]);
    mx = block_max(mx, red);
    float se = 0.f;
    for (int c = threadIdx.x; c < C; c += blockDim.x) se += __expf(x[c] - mx);
    se = block_sum(se, red);
    const float lse = mx + logf(se);
    if (lse_out && threadIdx.x == 0) lse_out[r] = lse;
    if (kind == 0) {
        const long long t = targets[r];
        const bool ign = (t == ignore_index);
        if (loss && threadIdx.x == 0) loss[r] = ign ? 0.f : lse - x[t];
        if (dlogits) {
            const float g = ign ? 0.f : dloss[r];
            for (int c = threadIdx.x; c < C; c += blockDim.x)
                dlogits[r * ld + c] = g * (__expf(x[c] - lse) - (c == t ? 1.f : 0.f));
        }
    } else {
        const float* t = soft + r * (long long)C;
        if (loss)
            for (int c = threadIdx.x; c < C; c += blockDim.x) {
                const float tv = t[c];
                // F.kl_div: target * (log(target) - input), 0 where target == 0
                loss[r * (long long)C + c] = tv > 0.f ? tv * (logf(tv) - (x[c] - lse)) : 0.f;
            }
        if (dlogits) {
            // d/dx_c sum_j dl_j * t_j * (log t_j - x_j + lse) = -dl_c t_c + softmax_c * sum_j dl_j t_j
            float s = 0.f;
            for (int c = threadIdx.x; c < C; c += blockDim.x) s += dloss[r * (long long)C + c] * t[c];
            s = block_sum(s, red);
            for (int c = threadIdx.x; c < C; c += blockDim.x)
                dlogits[r * ld + c] = -dloss[r * (long long)C + c] * t[c] + __expf(x[c] - lse) * s;
        }
    }
}

// Large-vocabulary cross entropy (the tied MLM decoder: rows x 250 002 fp32 logits, 1 MB per row).
// Forward: ONE pass over the row (online max / sum with 16-byte loads), writes loss and the row's log-sum-exp.
// Backward: one pass, reads the logits once more and writes d(logits) directly as the bf16, 8-element-pitched
// operand the dgrad / wgrad GEMMs consume (no fp32 d(logits) tensor, no separate cast kernel).
constexpr int CE_THREADS = 512;

__device__ __forceinline__ void online_add(float& m, float& s, float x) {
    if (x > m) { s = s * __expf(m - x) + 1.f; m = x; }
    else s += __expf(x - m);
}

__global__ void __launch_bounds__(CE_THREADS)
ce_loss_fwd_kernel(const float* __restrict__ logits, long long ld, int C, const long long* __restrict__ targets,
                   long long ignore_index, float* __restrict__ loss, float* __restrict__ lse_out) {
    __shared__ float red_m[CE_THREADS / 32], red_s[CE_THREADS / 32];
    const long long r = blockIdx.x;
    const float* x = logits + r * ld;
    float m = -INFINITY, s = 0.f;
    const int C4 = ((ld % 4) == 0 && (reinterpret_cast<uintptr_t>(logits) & 15) == 0) ? C / 4 : 0;
    for (int i = threadIdx.x; i < C4; i += CE_THREADS) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(x) + i);
        const float vm = fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w));
        if (vm > m) { s *= __expf(m - vm); m = vm; }
        s += __expf(v.x - m) + __expf(v.y - m) + __expf(v.z - m) + __expf(v.w - m);
    }
    for (int c = C4 * 4 + threadIdx.x; c < C; c += CE_THREADS) online_add(m, s, x[c]);
    // combine (m, s) pairs: warp, then block
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float m2 = __shfl_xor_sync(0xffffffffu, m, o), s2 = __shfl_xor_sync(0xffffffffu, s, o);
        const float mm = fmaxf(m, m2);
        s = (m == -INFINITY ? 0.f : s * __expf(m - mm)) + (m2 == -INFINITY ? 0.f : s2 * __expf(m2 - mm));
        m = mm;
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) { red_m[warp] = m; red_s[warp] = s; }
    __syncthreads();
    if (threadIdx.x == 0) {
        float M = red_m[0], S = red_s[0];
        for (int w = 1; w < CE_THREADS / 32; ++w) {
            const float mm = fmaxf(M, red_m[w]);
            S = (M == -INFINITY ? 0.f : S * __expf(M - mm)) + (red_m[w] == -INFINITY ? 0.f : red_s[w] * __expf(red_m[w] - mm));
            M = mm;
        }
        const float lse = M + logf(S);
        if (lse_out) lse_out[r] = lse;
        const long long t = targets[r];
        if (loss) loss[r] = (t == ignore_index) ? 0.f : lse - x[t];
    }
}

__global__ void __launch_bounds__(256)
ce_loss_bwd_kernel(const float* __restrict__ logits, long long ld, int C, const long long* __restrict__ targets,
                   long long ignore_index, const float* __restrict__ dloss, const float* __restrict__ lse,
                   bf16* __restrict__ dlogits, long long ld_out) {
    const long long r = blockIdx.y;
    const float* x = logits + r * ld;
    bf16* d = dlogits + r * ld_out;
    const long long t = targets[r];
    const float g = (t == ignore_index) ? 0.f : dloss[r];
    const float l = lse[r];
    const bool vec = (ld % 4) == 0 && (ld_out % 4) == 0 && (reinterpret_cast<uintptr_t>(logits) & 15) == 0 &&
                     (reinterpret_cast<uintptr_t>(dlogits) & 7) == 0;
    const int c0 = (blockIdx.x * 256 + threadIdx.x) * 4;
    if (c0 >= C) return;
    if (vec && c0 + 4 <= C) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(x + c0));
        float o[4] = {g * __expf(v.x - l), g * __expf(v.y - l), g * __expf(v.z - l), g * __expf(v.w - l)};
        if (t >= c0 && t < c0 + 4) o[t - c0] -= g;
        *reinterpret_cast<uint2*>(d + c0) = make_uint2(pack_bf16(o[0], o[1]), pack_bf16(o[2], o[3]));
    } else {
        for (int c = c0; c < C && c < c0 + 4; ++c)
            d[c] = __float2bfloat16(g * (__expf(x[c] - l) - (c == t ? 1.f : 0.f)));
    }
}

// ------------------------------------------------------------------------------------------------
// Fused cross entropy, second half: the decoder GEMM (gemm_tcgen05.cu, MODE 9) leaves one (max, sum exp(z - max))
// pair per row and 32-column chunk in stats[chunk][row]; merge them per row.  Level 1: a CTA takes 32 rows x 256
// chunks (lane = row, so every warp access is 256 contiguous bytes; warp = 32 of the chunks), level 2: one thread
// per row over the <= 31 partials, then lse and loss = lse - target logit (model/model.py:592-596).
// ------------------------------------------------------------------------------------------------
constexpr int CE_L1_CHUNKS = 256;

__device__ __forceinline__ void ce_merge(float& m, float& s, float m2, float s2) {
    const float mm = fmaxf(m, m2);
    s = (m == -INFINITY ? 0.f : s * __expf(m - mm)) + (m2 == -INFINITY ? 0.f : s2 * __expf(m2 - mm));
    m = mm;
}

__global__ void __launch_bounds__(256)
ce_stats_l1_kernel(const float2* __restrict__ stats, long long ld, int n_chunks, long long rows, float2* __restrict__ part) {
    __shared__ float2 red[8][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long row = (long long)blockIdx.y * 32 + lane;
    const int c0 = blockIdx.x * CE_L1_CHUNKS + warp * 32;
    float m = -INFINITY, s = 0.f;
    if (row < rows) {
        const int c1 = min(c0 + 32, n_chunks);
#pragma unroll 4
        for (int c = c0; c < c1; ++c) {
            const float2 v = __ldg(stats + (long long)c * ld + row);
            ce_merge(m, s, v.x, v.y);
        }
    }
    red[warp][lane] = make_float2(m, s);
    __syncthreads();
    if (warp == 0 && row < rows) {
#pragma unroll
        for (int w = 1; w < 8; ++w) ce_merge(m, s, red[w][lane].x, red[w][lane].y);
        part[(long long)blockIdx.x * rows + row] = make_float2(m, s);
    }
}

__global__ void __launch_bounds__(128)
ce_stats_l2_kernel(const float2* __restrict__ part, int n_part, long long rows, const float* __restrict__ tgt,
                   const long long* __restrict__ targets, long long ignore_index, float* __restrict__ loss,
                   float* __restrict__ lse_out) {
    const long long row = (long long)blockIdx.x * 128 + threadIdx.x;
    if (row >= rows) return;
    float m = -INFINITY, s = 0.f;
    for (int k = 0; k < n_part; ++k) {
        const float2 v = part[(long long)k * rows + row];
        ce_merge(m, s, v.x, v.y);
    }
    const float lse = m + logf(s);
    lse_out[row] = lse;
    if (loss) loss[row] = targets[row] == ignore_index ? 0.f : lse - tgt[row];
}

// d(logits) over the bf16 logits, in place (8 columns per thread)
__global__ void __launch_bounds__(256)
ce_bwd_inplace_kernel(bf16* __restrict__ z, long long ld, int C, const long long* __restrict__ targets,
                      long long ignore_index, const float* __restrict__ dloss, const float* __restrict__ lse) {
    const long long r = blockIdx.y;
    bf16* x = z + r * ld;
    const long long t = targets[r];
    const float g = (t == ignore_index) ? 0.f : dloss[r];
    const float l = lse[r];
    const int c0 = (blockIdx.x * 256 + threadIdx.x) * 8;
    if (c0 >= C) return;
    if (c0 + 8 <= C) {
        float v[8];
        load8_bf16(x + c0, v);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = g * (__expf(v[j] - l) - ((long long)(c0 + j) == t ? 1.f : 0.f));
        store8_bf16(x + c0, v);
    } else {
        for (int c = c0; c < C; ++c)
            x[c] = __float2bfloat16(g * (__expf(__bfloat162float(x[c]) - l) - ((long long)c == t ? 1.f : 0.f)));
    }
}

// ------------------------------------------------------------------------------------------------
// order-exact masked-row compaction (hidden[mask], model/model.py:653-657), no host sync:
//   pass 1 (one CTA): exclusive scan of the mask in row-major (b, j) order -> index list + count
//   pass 2: gather rows / scatter-add gradient rows
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024)
mask_scan_kernel(const unsigned char* __restrict__ mask, long long n, int* __restrict__ index, int* __restrict__ count,
                 int capacity) {
    __shared__ int warp_tot[32];
    __shared__ int carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (long long base = 0; base < n; base += 1024) {
        const long long i = base + threadIdx.x;
        const int f = (i < n && mask[i]) ? 1 : 0;
        const unsigned bal = __ballot_sync(0xffffffffu, f);
        const int pre = __popc(bal & ((1u << lane) - 1));
        if (lane == 0) warp_tot[warp] = __popc(bal);
        __syncthreads();
        int off = carry;
        for (int w = 0; w < warp; ++w) off += warp_tot[w];
        if (f && off + pre < capacity) index[off + pre] = (int)i;
        __syncthreads();
        if (threadIdx.x == 0) {
            int t = 0;
            for (int w = 0; w < 32; ++w) t += warp_tot[w];
            carry += t;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) *count = carry;
}

// out[i,:] = src[index[i] mapped through (b, j) -> b*src_S + j, :]  (rows of 768 bf16); i < *count
__global__ void __launch_bounds__(256)
gather_rows_kernel(const bf16* __restrict__ src, const int* __restrict__ index, const int* __restrict__ count,
                   int mask_L, int src_S, bf16* __restrict__ out, int capacity) {
    const int lane = threadIdx.x & 31;
    const int i = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int n = min(*count, capacity);
    if (i >= n) return;
    const int flat = index[i];
    const long long row = (long long)(flat / mask_L) * src_S + (flat % mask_L);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const int col = c * 256 + lane * 8;
        *reinterpret_cast<uint4*>(out + (long long)i * HID + col) =
            *reinterpret_cast<const uint4*>(src + row * HID + col);
    }
}

// dsrc[row(index[i]), :] += dout[i, :]   (rows are unique, so plain read-modify-write is race free)
__global__ void __launch_bounds__(256)
scatter_rows_add_kernel(const bf16* __restrict__ dout, const int* __restrict__ index, const int* __restrict__ count,
                        int mask_L, int src_S, bf16* __restrict__ dsrc, int capacity) {
    const int lane = threadIdx.x & 31;
    const int i = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int n = min(*count, capacity);
    if (i >= n) return;
    const int flat = index[i];
    const long long row = (long long)(flat / mask_L) * src_S + (flat % mask_L);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const int col = c * 256 + lane * 8;
        float a[8], b[8];
        load8_bf16(dout + (long long)i * HID + col, a);
        load8_bf16(dsrc + row * HID + col, b);
#pragma unroll
        for (int k = 0; k < 8; ++k) a[k] += b[k];
        store8_bf16(dsrc + row * HID + col, a);
    }
}

// elementwise helpers ------------------------------------------------------------------------------
// mse 'none': loss = (pred - tgt)^2 ; dpred = 2 (pred - tgt) dloss
__global__ void mse_kernel(const float* __restrict__ pred, const float* __restrict__ tgt, float* __restrict__ loss,
                           const float* __restrict__ dloss, float* __restrict__ dpred, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float d = pred[i] - tgt[i];
    if (loss) loss[i] = d * d;
    if (dpred) dpred[i] = 2.f * d * dloss[i];
}
// dz = dy * gelu'(pre), all bf16 (head transforms)
__global__ void dgelu_bf16_kernel(const bf16* __restrict__ dy, const bf16* __restrict__ pre, bf16* __restrict__ out,
                                  long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = __float2bfloat16(__bfloat162float(dy[i]) * gelu_erf_grad(__bfloat162float(pre[i])));
}
__global__ void f32_to_bf16_strided_kernel(const float* __restrict__ x, long long ldx, bf16* __restrict__ y,
                                           long long ldy, long long rows, int cols) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * cols) return;
    const long long r = i / cols;
    const int c = (int)(i % cols);
    y[r * ldy + c] = __float2bfloat16(x[r * ldx + c]);
}

}  // namespace
}  // namespace uc2

using namespace uc2;

extern "C" UC2_API int uc2_narrow_linear_fwd(const float* x, long long ldx, const float* W, const float* b, float* out,
                                             int M, int N, int K, void* stream) {
    if (int rc = require_sm100()) return rc;
    UC2_REQUIRE(x && W && out && M > 0 && N > 0 && N <= 8 && K % 128 == 0 && ldx % 4 == 0, UC2_ERR_ARG,
                "narrow_linear_fwd: bad args (N<=8, K%%128==0)");
    narrow_linear_fwd_kernel<<<(M + 7) / 8, 256, 0, (cudaStream_t)stream>>>(x, ldx, W, b, out, M, N, K);
    return check_last("narrow_linear_fwd_kernel");
}

extern "C" UC2_API int uc2_narrow_linear_bwd(const float* x, long long ldx, const float* W, const float* dy, float* dx,
                                             long long lddx, float* dW, float* db, int M, int N, int K, void* stream) {
    if (int rc = require_sm100()) return rc;
    UC2_REQUIRE(x && W && dy && dW && M > 0 && N > 0 && N <= 8 && K >= N, UC2_ERR_ARG, "narrow_linear_bwd: bad args");
    narrow_linear_bwd_kernel<<<(K + 127) / 128, 128, 0, (cudaStream_t)stream>>>(x, ldx, W, dy, dx, lddx, dW, db, M, N, K);
    return check_last("narrow_linear_bwd_kernel");
}

extern "C" UC2_API int uc2_tanh_bwd(const float* y, const float* dy, void* dpre_bf16, long long n, void* stream) {
    if (int rc = require_sm100()) return rc;
    UC2_REQUIRE(y && dy && dpre_bf16 && n > 0, UC2_ERR_ARG, "tanh_bwd: bad args");
    tanh_bwd_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(y, dy, (bf16*)dpre_bf16, n);
    return check_last("tanh_bwd_kernel");
}

extern "C" UC2_API int uc2_rank_loss_fwd(const float* scores, float* loss, int groups, int sample_size, float margin,
                                         void* stream) {
    if (int rc = require_sm100()) return rc;
    UC2_REQUIRE(scores && loss && groups > 0 && sample_size > 1, UC2_ERR_ARG, "rank_loss_fwd: bad args");
    const int n = groups * (sample_size - 1);
    rank_loss_fwd_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(scores, loss, groups, sample_size, margin);
    return check_last("rank_loss_fwd_kernel");
}

extern "C" UC2_API int uc2_rank_loss_bwd(const float* scores, const float* dloss, float* dscores, int groups,
                                         int sample_size, float margin, void* stream) {
    if (int rc = require_sm100()) return rc;
    UC2_REQUIRE(scores && dloss && dscores && groups > 0 && sample_size > 1, UC2_ERR_ARG, "rank_loss_bwd: bad args");
    rank_loss_bwd_kernel<<<(groups + 127) / 128, 128, 0, (cudaStream_t)stream>>>(scores, dloss, dscores, groups,
                                                                                 sample_size, margin);
    return check_last("rank_loss_bwd_kernel");
}

extern "C" UC2_API int uc2_softmax_loss(const float* logits, long long ld, long long rows, int C, int kind,
                                        const long long* targets, long long ignore_index, const float* soft_targets,
                                        float* loss, const float* dloss, float* dlogits, float* lse_out, void* stream) {
    if (int rc = require_sm100()) return rc;
    UC2_REQUIRE(logits && rows >= 0 && C > 0 && ld >= C, UC2_ERR_ARG, "softmax_loss: bad args");
    UC2_REQUIRE((kind == 0 && targets) || (kind == 1 && soft_targets), UC2_ERR_ARG, "softmax_loss: targets missing");
    UC2_REQUIRE(!dlogits || dloss, UC2_ERR_ARG, "softmax_loss: backward needs dloss");
    if (rows == 0) return UC2_OK;
    softmax_loss_kernel<<<(unsigned)rows, 256, 0, (cudaStream_t)stream>>>(logits, ld, C, kind, targets, ignore_index,
                                                                         soft_targets, loss, dloss, dlogits, lse_out);
    return check_last("softmax_loss_kernel");
}

extern "C" UC2_API int uc2_ce_loss_fwd(const float* logits, long long ld, long long rows, int C,
                                       const long long* targets, long long ignore_index, float* loss, float* lse,
                                       void* stream) {
    if (int rc = require_sm100()) return rc;
    UC2_REQUIRE(logits && targets && lse && rows >= 0 && C > 0 && ld >= C, UC2_ERR_ARG, "ce_loss_fwd: bad args");
    if (rows == 0) return UC2_OK;
    ce_loss_fwd_kernel<<<(unsigned)rows, CE_THREADS, 0, (cudaStream_t)stream>>>(logits, ld, C, targets, ignore_index,
                                                                              loss, lse);
    return check_last("ce_loss_fwd_kernel");
}

extern "C" UC2_API int uc2_ce_loss_bwd_bf16(const float* logits, long long ld, long long rows, int C,
                                            const long long* targets, long long ignore_index, const float* dloss,
                                            const float* lse, void* dlogits_bf16, long long ld_out, void* stream) {
    if (int rc = require_sm100()) return rc;
    UC2_REQUIRE(logits && targets && dloss && lse && dlogits_bf16 && rows >= 0 && C > 0 && ld >= C && ld_out >= C,
                UC2_ERR_ARG, "ce_loss_bwd: bad args");
    UC2_REQUIRE(rows < 65536, UC2_ERR_UNSUPPORTED, "ce_loss_bwd: at most 65535 rows per call");
    if (rows == 0) return UC2_OK;
    ce_loss_bwd_kernel<<<dim3((C + 1023) / 1024, (unsigned)rows), 256, 0, (cudaStream_t)stream>>>(
        logits, ld, C, targets, ignore_index, dloss, lse, (bf16*)dlogits_bf16, ld_out);
    return check_last("ce_loss_bwd_kernel");
}

extern "C" UC2_API int uc2_ce_stats_reduce(const float* stats, long long ld_ce, int n_chunks, long long rows,
                                           const float* tgt, const long long* targets, long long ignore_index,
                                           float* part, float* loss, float* lse, void* stream) {
    if (int rc = require_sm100()) return rc;
    UC2_REQUIRE(stats && tgt && targets && part && lse && n_chunks > 0 && rows >= 0 && ld_ce >= rows &&
                    (reinterpret_cast<uintptr_t>(stats) & 7) == 0 && (reinterpret_cast<uintptr_t>(part) & 7) == 0,
                UC2_ERR_ARG, "ce_stats_reduce: bad args");
    if (rows == 0) return UC2_OK;
    const int n_part = (n_chunks + CE_L1_CHUNKS - 1) / CE_L1_CHUNKS;
    UC2_REQUIRE((rows + 31) / 32 < 65536, UC2_ERR_UNSUPPORTED, "ce_stats_reduce: too many rows");
    ce_stats_l1_kernel<<<dim3(n_part, (unsigned)((rows + 31) / 32)), 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const float2*>(stats), ld_ce, n_chunks, rows, reinterpret_cast<float2*>(part));
    if (int rc = check_last("ce_stats_l1_kernel")) return rc;
    ce_stats_l2_kernel<<<(unsigned)((rows + 127) / 128), 128, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const float2*>(part), n_part, rows, tgt, targets, ignore_index, loss, lse);
    return check_last("ce_stats_l2_kernel");
}

extern "C" UC2_API int uc2_ce_bwd_inplace_bf16(void* logits_bf16, long long ld, long long rows, int C,
                                               const long long* targets, long long ignore_index, const float* dloss,
                                               const float* lse, void* stream) {
    if (int rc = require_sm100()) return rc;
    UC2_REQUIRE(logits_bf16 && targets && dloss && lse && rows >= 0 && C > 0 && ld >= C && ld % 8 == 0 &&
                    aligned16(logits_bf16),
                UC2_ERR_ARG, "ce_bwd_inplace: bad args (ld must be a multiple of 8, base 16-byte aligned)");
    UC2_REQUIRE(rows < 65536, UC2_ERR_UNSUPPORTED, "ce_bwd_inplace: at most 65535 rows per call");
    if (rows == 0) return UC2_OK;
    ce_bwd_inplace_kernel<<<dim3((C + 2047) / 2048, (unsigned)rows), 256, 0, (cudaStream_t)stream>>>(
        (bf16*)logits_bf16, ld, C, targets, ignore_index, dloss, lse);
    return check_last("ce_bwd_inplace_kernel");
}

extern "C" UC2_API int uc2_mask_scan(const unsigned char* mask, long long n, int* index, int* count, int capacity,
                                     void* stream) {
    if (int rc = require_sm100()) return rc;
    UC2_REQUIRE(mask && index && count && n > 0 && n < (1LL << 31) && capacity > 0, UC2_ERR_ARG, "mask_scan: bad args");
    mask_scan_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(mask, n, index, count, capacity);
    return check_last("mask_scan_kernel");
}

extern "C" UC2_API int uc2_gather_rows(const void* src, const int* index, const int* count, int mask_L, int src_S,
                                       void* out, int capacity, void* stream) {
    if (int rc = require_sm100()) return rc;
    UC2_REQUIRE(src && index && count && out && capacity > 0 && mask_L > 0 && src_S >= mask_L, UC2_ERR_ARG,
                "gather_rows: bad args");
    gather_rows_kernel<<<(capacity + 7) / 8, 256, 0, (cudaStream_t)stream>>>((const bf16*)src, index, count, mask_L,
                                                                            src_S, (bf16*)out, capacity);
    return check_last("gather_rows_kernel");
}

extern "C" UC2_API int uc2_scatter_rows_add(const void* dout, const int* index, const int* count, int mask_L,
                                            int src_S, void* dsrc, int capacity, void* stream) {
    if (int rc = require_sm100()) return rc;
    UC2_REQUIRE(dout && index && count && dsrc && capacity > 0 && mask_L > 0 && src_S >= mask_L, UC2_ERR_ARG,
                "scatter_rows_add: bad args");
    scatter_rows_add_kernel<<<(capacity + 7) / 8, 256, 0, (cudaStream_t)stream>>>((const bf16*)dout, index, count,
                                                                                 mask_L, src_S, (bf16*)dsrc, capacity);
    return check_last("scatter_rows_add_kernel");
}

extern "C" UC2_API int uc2_mse(const float* pred, const float* tgt, float* loss, const float* dloss, float* dpred,
                               long long n, void* stream) {
    if (int rc = require_sm100()) return rc;
    UC2_REQUIRE(pred && tgt && n >= 0 && (!dpred || dloss), UC2_ERR_ARG, "mse: bad args");
    if (n == 0) return UC2_OK;
    mse_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(pred, tgt, loss, dloss, dpred, n);
    return check_last("mse_kernel");
}

extern "C" UC2_API int uc2_f32_to_bf16_2d(const float* x, long long ldx, void* y, long long ldy, long long rows,
                                          int cols, void* stream) {
    if (int rc = require_sm100()) return rc;
    UC2_REQUIRE(x && y && rows > 0 && cols > 0, UC2_ERR_ARG, "f32_to_bf16_2d: bad args");
    const long long n = rows * cols;
    f32_to_bf16_strided_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(x, ldx, (bf16*)y, ldy, rows,
                                                                                            cols);
    return check_last("f32_to_bf16_strided_kernel");
}

extern "C" UC2_API int uc2_dgelu_bf16(const void* dy, const void* pre, void* out, long long n, void* stream) {
    if (int rc = require_sm100()) return rc;
    UC2_REQUIRE(dy && pre && out && n > 0, UC2_ERR_ARG, "dgelu_bf16: bad args");
    dgelu_bf16_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>((const bf16*)dy, (const bf16*)pre,
                                                                                   (bf16*)out, n);
    return check_last("dgelu_bf16_kernel");
}
