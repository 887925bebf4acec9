// Flat-arena optimiser kernels: fp32 -> bf16 shadow cast, global gradient norm, and a single-launch
// multi-tensor AdamW that follows optim/adamw.py:40-103 exactly (Adam update with bias correction folded
// into the step size, eps added outside the sqrt, THEN decoupled decay on the updated weight), fused with
// clip_grad_norm_ (pretrain.py:610), the bf16 shadow refresh and zero_grad.
// HBM-bound: 16 B read + 18 B written per parameter.
#include "common.cuh"

namespace uc2 {
namespace {

constexpr int OPT_THREADS = 256;
constexpr int CHUNK = 8192;          // elements per CTA work item (32 per thread)

__global__ void __launch_bounds__(OPT_THREADS)
cast_f32_bf16_kernel(const float* __restrict__ src, bf16* __restrict__ dst, long long n) {
    const long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 8;
    if (i + 8 <= n) {
        float v[8];
        load8_f32(src + i, v);
        store8_bf16(dst + i, v);
    } else {
        for (long long k = i; k < n; ++k) dst[k] = __float2bfloat16(src[k]);
    }
}

// This thread's share of the sum of squares of one <= CHUNK-element gradient slice (fixed order and rounding: the
// eager and the row-sparse norm kernels must produce the same per-chunk partial sums)
__device__ __forceinline__ float chunk_sqsum(const float* __restrict__ g, int n) {
    float acc = 0.f;
    for (int i = threadIdx.x * 4; i < n; i += OPT_THREADS * 4) {
        if (i + 4 <= n) {
            const float4 v = *reinterpret_cast<const float4*>(g + i);
            acc = __fadd_rn(acc, __fmaf_rn(v.w, v.w, __fmaf_rn(v.z, v.z, __fmaf_rn(v.y, v.y, __fmul_rn(v.x, v.x)))));
        } else {
            for (int k = i; k < n; ++k) acc = __fmaf_rn(g[k], g[k], acc);
        }
    }
    return acc;
}

// block-wide sum of the per-thread shares of ONE chunk, added to *out in double (an all-zero chunk adds nothing)
__device__ __forceinline__ void chunk_sqsum_commit(float acc, float* red, double* out) {
    acc = warp_sum(acc);
    __syncthreads();                                 // red[] of the previous chunk has been consumed
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int w = 0; w < OPT_THREADS / 32; ++w) s += red[w];
        if (s != 0.0) atomicAdd(out, s);
    }
}

// sum of squares of the gradients of ACTIVE tensors (act_step >= 0) into *out (double), one partial per chunk
__global__ void __launch_bounds__(OPT_THREADS)
grad_sqnorm_kernel(const float* __restrict__ grad, const uc2_opt_chunk* __restrict__ chunks, int n_chunks,
                   const int* __restrict__ act_step, double* __restrict__ out) {
    __shared__ float red[OPT_THREADS / 32];
    for (int c = blockIdx.x; c < n_chunks; c += gridDim.x) {
        const uc2_opt_chunk ch = chunks[c];
        if (act_step[ch.tensor] < 0) continue;
        chunk_sqsum_commit(chunk_sqsum(grad + ch.offset, ch.n), red, out);
    }
}

// One element, one step of optim/adamw.py:77-101 (g already scaled by the clip coefficient).  Shared by the eager and
// the deferred kernels: the deferred replay has to reproduce the eager result bit for bit.
__device__ __forceinline__ void adamw_update(float& p, float& m, float& v, float g, float beta1, float beta2, float eps,
                                             float step_size, float decay, bool decay_on) {
    // every rounding spelled out: the compiler must not contract these differently in different kernels
    m = __fmaf_rn(1.0f - beta1, g, __fmul_rn(m, beta1));
    v = __fmaf_rn(__fmul_rn(1.0f - beta2, g), g, __fmul_rn(v, beta2));
    const float denom = __fadd_rn(__fsqrt_rn(v), eps);
    p = __fmaf_rn(-step_size, __fdiv_rn(m, denom), p);
    if (decay_on) p = __fmaf_rn(-decay, p, p);
}

// step size of tensor-step `step` with the bias correction folded in (adamw.py:87-92)
__device__ __forceinline__ float adamw_step_size(float lr, float beta1, float beta2, int step, int correct_bias) {
    float step_size = lr;
    if (correct_bias) {
        const double bc1 = 1.0 - pow((double)beta1, (double)step);
        const double bc2 = 1.0 - pow((double)beta2, (double)step);
        step_size = (float)((double)lr * sqrt(bc2) / bc1);
    }
    return step_size;
}

__device__ __forceinline__ float clip_coef(float max_grad_norm, const double* sqnorm) {
    // torch.nn.utils.clip_grad_norm_: max_norm / (total_norm + 1e-6), capped at 1
    if (max_grad_norm > 0.f && sqnorm) {
        const float total = (float)sqrt(*sqnorm);
        return fminf(max_grad_norm / (total + 1e-6f), 1.0f);
    }
    return 1.0f;
}

__global__ void __launch_bounds__(OPT_THREADS)
adamw_kernel(float* __restrict__ param, float* __restrict__ grad, float* __restrict__ exp_avg,
             float* __restrict__ exp_avg_sq, bf16* __restrict__ shadow, const uc2_opt_chunk* __restrict__ chunks,
             int n_chunks, const int* __restrict__ act_step, const int* __restrict__ group_of,
             const uc2_adamw_hyper h, const double* __restrict__ sqnorm) {
    const float clip = clip_coef(h.max_grad_norm, sqnorm);
    for (int c = blockIdx.x; c < n_chunks; c += gridDim.x) {
        const uc2_opt_chunk ch = chunks[c];
        const int a = act_step[ch.tensor];
        if (a < 0) continue;                           // p.grad is None -> `continue` (adamw.py:52-53)
        const int grp = group_of[ch.tensor];
        const float lr = h.lr[grp], wd = h.weight_decay[grp];
        const int step = h.global_step - a + 1;        // state['step'] of this tensor
        const float step_size = adamw_step_size(lr, h.beta1, h.beta2, step, h.correct_bias);
        const float decay = lr * wd;
        float* p = param + ch.offset;
        float* g = grad + ch.offset;
        float* m = exp_avg + ch.offset;
        float* v = exp_avg_sq + ch.offset;
        bf16* sh = shadow ? shadow + ch.offset : nullptr;
        for (int i = threadIdx.x * 4; i < ch.n; i += OPT_THREADS * 4) {
            float pv[4], gv[4], mv[4], vv[4];
            const bool full = i + 4 <= ch.n;
            if (full) {
                *reinterpret_cast<float4*>(pv) = *reinterpret_cast<const float4*>(p + i);
                *reinterpret_cast<float4*>(gv) = *reinterpret_cast<const float4*>(g + i);
                *reinterpret_cast<float4*>(mv) = *reinterpret_cast<const float4*>(m + i);
                *reinterpret_cast<float4*>(vv) = *reinterpret_cast<const float4*>(v + i);
            } else {
                for (int k = 0; k < 4; ++k) {
                    const bool ok = i + k < ch.n;
                    pv[k] = ok ? p[i + k] : 0.f; gv[k] = ok ? g[i + k] : 0.f;
                    mv[k] = ok ? m[i + k] : 0.f; vv[k] = ok ? v[i + k] : 0.f;
                }
            }
#pragma unroll
            for (int k = 0; k < 4; ++k)
                adamw_update(pv[k], mv[k], vv[k], gv[k] * clip, h.beta1, h.beta2, h.eps, step_size, decay, wd > 0.f);
            if (full) {
                *reinterpret_cast<float4*>(p + i) = *reinterpret_cast<float4*>(pv);
                *reinterpret_cast<float4*>(m + i) = *reinterpret_cast<float4*>(mv);
                *reinterpret_cast<float4*>(v + i) = *reinterpret_cast<float4*>(vv);
                if (h.zero_grad) *reinterpret_cast<float4*>(g + i) = make_float4(0.f, 0.f, 0.f, 0.f);
                if (sh) *reinterpret_cast<uint2*>(sh + i) = make_uint2(pack_bf16(pv[0], pv[1]), pack_bf16(pv[2], pv[3]));
            } else {
                for (int k = 0; k < 4 && i + k < ch.n; ++k) {
                    p[i + k] = pv[k]; m[i + k] = mv[k]; v[i + k] = vv[k];
                    if (h.zero_grad) g[i + k] = 0.f;
                    if (sh) sh[i + k] = __float2bfloat16(pv[k]);
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Deferred AdamW for a row-sparse table (include/uc2_b200.h, uc2_lazy_table)
// ------------------------------------------------------------------------------------------------
struct LazyScalars { float step_size, decay; };

__device__ __forceinline__ LazyScalars lazy_hist(const uc2_lazy_table& t, int step) {
    const float2 e = __ldg(reinterpret_cast<const float2*>(t.hist) + (step % t.hist_len));
    return LazyScalars{e.x, e.y};
}

// One CTA per entry of row_ids, one float4 per thread.  GRAD: apply step `step` with the row's gradient after the replay
// (else: replay only, through `step`).  Duplicate ids: the first CTA to raise row_step to `step` owns the row.
template <bool GRAD>
__global__ void __launch_bounds__(256)
adamw_lazy_rows_kernel(const uc2_lazy_table t, const long long* __restrict__ row_ids, int step, float max_grad_norm,
                       const double* __restrict__ sqnorm) {
    __shared__ int s_old;
    const long long r = row_ids ? row_ids[blockIdx.x] : static_cast<long long>(blockIdx.x);
    if (r < 0 || r >= t.n_rows) return;
    if (threadIdx.x == 0) s_old = atomicMax(t.row_step + r, step);
    __syncthreads();
    const int old = s_old;
    if (old >= step) return;                         // up to date already, or another CTA of this launch has the row
    const float clip = GRAD ? clip_coef(max_grad_norm, sqnorm) : 1.0f;
    const bool decay_on = t.decay_on != 0;
    const long long base = t.table_off + r * t.width;
    bf16* shadow = static_cast<bf16*>(t.shadow_bf16);
    const int last_replayed = GRAD ? step - 1 : step;
    for (int i = threadIdx.x * 4; i < t.width; i += blockDim.x * 4) {
        float pv[4], mv[4], vv[4];
        *reinterpret_cast<float4*>(pv) = *reinterpret_cast<const float4*>(t.param + base + i);
        *reinterpret_cast<float4*>(mv) = *reinterpret_cast<const float4*>(t.exp_avg + base + i);
        *reinterpret_cast<float4*>(vv) = *reinterpret_cast<const float4*>(t.exp_avg_sq + base + i);
        for (int k = old + 1; k <= last_replayed; ++k) {           // the steps this row sat out: gradient 0
            const LazyScalars h = lazy_hist(t, k);
#pragma unroll
            for (int e = 0; e < 4; ++e)
                adamw_update(pv[e], mv[e], vv[e], 0.f, t.beta1, t.beta2, t.eps, h.step_size, h.decay, decay_on);
        }
        if (GRAD) {
            float gv[4];
            *reinterpret_cast<float4*>(gv) = *reinterpret_cast<const float4*>(t.grad + base + i);
            const LazyScalars h = lazy_hist(t, step);
#pragma unroll
            for (int e = 0; e < 4; ++e)
                adamw_update(pv[e], mv[e], vv[e], gv[e] * clip, t.beta1, t.beta2, t.eps, h.step_size, h.decay, decay_on);
            *reinterpret_cast<float4*>(t.grad + base + i) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        *reinterpret_cast<float4*>(t.param + base + i) = *reinterpret_cast<float4*>(pv);
        *reinterpret_cast<float4*>(t.exp_avg + base + i) = *reinterpret_cast<float4*>(mv);
        *reinterpret_cast<float4*>(t.exp_avg_sq + base + i) = *reinterpret_cast<float4*>(vv);
        if (shadow) *reinterpret_cast<uint2*>(shadow + base + i) = make_uint2(pack_bf16(pv[0], pv[1]), pack_bf16(pv[2], pv[3]));
    }
}

// the ring entry of step `step`, derived by the same device code as the eager kernel's scalars
__global__ void adamw_lazy_note_kernel(const uc2_lazy_table t, int step, int tensor_step, float lr, float wd, int correct_bias) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        reinterpret_cast<float2*>(t.hist)[step % t.hist_len] =
            make_float2(adamw_step_size(lr, t.beta1, t.beta2, tensor_step, correct_bias), lr * wd);
    }
}

// Norm contribution of the table on a row-sparse step: the chunks (the eager kernel's CHUNK-element slices of the table)
// that hold a listed row, every distinct chunk once, with the eager kernel's own per-chunk arithmetic; every other
// chunk of the table is all zero and would add nothing.  Two CTAs per id: the chunk of the row's first / last element.
__global__ void __launch_bounds__(OPT_THREADS)
grad_sqnorm_rows_kernel(const uc2_lazy_table t, const long long* __restrict__ row_ids, int mark, double* __restrict__ out) {
    __shared__ float red[OPT_THREADS / 32];
    __shared__ int s_old;
    const long long r = row_ids[blockIdx.x >> 1];
    if (r < 0 || r >= t.n_rows) return;
    const long long total = (long long)t.n_rows * t.width;
    const long long c = (r * t.width + ((blockIdx.x & 1) ? t.width - 1 : 0)) / CHUNK;
    if (threadIdx.x == 0) s_old = atomicExch(t.row_seen + c, mark);
    __syncthreads();
    if (s_old == mark) return;
    const long long start = c * CHUNK;
    const int n = (int)(total - start < CHUNK ? total - start : CHUNK);
    chunk_sqsum_commit(chunk_sqsum(t.grad + t.table_off + start, n), red, out);
}

int lazy_check(const uc2_lazy_table* t, const char* who) {
    UC2_REQUIRE(t && t->param && t->grad && t->exp_avg && t->exp_avg_sq && t->row_step && t->row_seen && t->hist,
                UC2_ERR_ARG, "%s: null table field", who);
    UC2_REQUIRE(t->n_rows > 0 && t->width > 0 && t->width % 4 == 0 && t->table_off % 4 == 0 && t->hist_len >= 2 &&
                    aligned16(t->param) && aligned16(t->grad) && aligned16(t->exp_avg) && aligned16(t->exp_avg_sq) &&
                    (reinterpret_cast<uintptr_t>(t->hist) & 7) == 0,
                UC2_ERR_ARG, "%s: bad table geometry", who);
    return UC2_OK;
}

}  // namespace
}  // namespace uc2

using namespace uc2;

extern "C" UC2_API int uc2_adamw_lazy_note(const uc2_lazy_table* t, int step, int first_step, float lr,
                                           float weight_decay, int correct_bias, void* stream) {
    if (int rc = require_sm100()) return rc;
    if (int rc = lazy_check(t, "adamw_lazy_note")) return rc;
    UC2_REQUIRE(step >= 1 && first_step >= 1 && first_step <= step, UC2_ERR_ARG, "adamw_lazy_note: bad step");
    adamw_lazy_note_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(*t, step, step - first_step + 1, lr, weight_decay, correct_bias);
    return check_last("adamw_lazy_note_kernel");
}

extern "C" UC2_API int uc2_adamw_lazy_rows(const uc2_lazy_table* t, const long long* row_ids, long long n_ids, int step,
                                           int first_step, float lr, float weight_decay, int correct_bias,
                                           float max_grad_norm, const double* sqnorm, void* stream) {
    if (int rc = uc2_adamw_lazy_note(t, step, first_step, lr, weight_decay, correct_bias, stream)) return rc;
    if (n_ids <= 0) return UC2_OK;
    UC2_REQUIRE(row_ids, UC2_ERR_ARG, "adamw_lazy_rows: null row_ids");
    UC2_REQUIRE((weight_decay > 0.f) == (t->decay_on != 0), UC2_ERR_ARG, "adamw_lazy_rows: decay_on does not match weight_decay");
    // the kernel reads this step's scalars from the ring entry the note kernel has just written (same stream)
    adamw_lazy_rows_kernel<true><<<(unsigned)n_ids, 192, 0, (cudaStream_t)stream>>>(*t, row_ids, step, max_grad_norm, sqnorm);
    return check_last("adamw_lazy_rows_kernel");
}

extern "C" UC2_API int uc2_adamw_lazy_catchup(const uc2_lazy_table* t, const long long* row_ids, long long n_ids, int upto,
                                              void* stream) {
    if (int rc = require_sm100()) return rc;
    if (int rc = lazy_check(t, "adamw_lazy_catchup")) return rc;
    const long long n = row_ids ? n_ids : t->n_rows;
    if (n <= 0 || upto < 1) return UC2_OK;
    adamw_lazy_rows_kernel<false><<<(unsigned)n, 192, 0, (cudaStream_t)stream>>>(*t, row_ids, upto, 0.f, nullptr);
    return check_last("adamw_lazy_rows_kernel");
}

extern "C" UC2_API int uc2_grad_sqnorm_rows(const uc2_lazy_table* t, const long long* row_ids, long long n_ids, int mark,
                                            double* out, void* stream) {
    if (int rc = require_sm100()) return rc;
    if (int rc = lazy_check(t, "grad_sqnorm_rows")) return rc;
    UC2_REQUIRE(out && (row_ids || n_ids == 0), UC2_ERR_ARG, "grad_sqnorm_rows: bad args");
    if (n_ids <= 0) return UC2_OK;
    grad_sqnorm_rows_kernel<<<(unsigned)(2 * n_ids), OPT_THREADS, 0, (cudaStream_t)stream>>>(*t, row_ids, mark, out);
    return check_last("grad_sqnorm_rows_kernel");
}

extern "C" UC2_API int uc2_cast_f32_bf16(const float* src, void* dst, long long n, void* stream) {
    if (int rc = require_sm100()) return rc;
    UC2_REQUIRE(src && dst && n > 0, UC2_ERR_ARG, "cast_f32_bf16: bad args");
    UC2_REQUIRE(aligned16(src) && aligned16(dst), UC2_ERR_ARG, "cast_f32_bf16: pointers must be 16-byte aligned");
    const long long threads = (n + 7) / 8;
    cast_f32_bf16_kernel<<<(unsigned)((threads + OPT_THREADS - 1) / OPT_THREADS), OPT_THREADS, 0,
                           (cudaStream_t)stream>>>(src, (bf16*)dst, n);
    return check_last("cast_f32_bf16_kernel");
}

extern "C" UC2_API int uc2_grad_sqnorm(const float* grad, const uc2_opt_chunk* chunks, int n_chunks,
                                       const int* act_step, double* out, void* stream) {
    if (int rc = require_sm100()) return rc;
    UC2_REQUIRE(grad && chunks && act_step && out && n_chunks > 0, UC2_ERR_ARG, "grad_sqnorm: bad args");
    UC2_CUDA(cudaMemsetAsync(out, 0, sizeof(double), (cudaStream_t)stream));
    const int blocks = n_chunks < 8 * num_sms() ? n_chunks : 8 * num_sms();
    grad_sqnorm_kernel<<<blocks, OPT_THREADS, 0, (cudaStream_t)stream>>>(grad, chunks, n_chunks, act_step, out);
    return check_last("grad_sqnorm_kernel");
}

extern "C" UC2_API int uc2_adamw_step(float* param, float* grad, float* exp_avg, float* exp_avg_sq, void* shadow_bf16,
                                      const uc2_opt_chunk* chunks, int n_chunks, const int* act_step,
                                      const int* group_of, const uc2_adamw_hyper* hyper, const double* sqnorm,
                                      void* stream) {
    if (int rc = require_sm100()) return rc;
    UC2_REQUIRE(param && grad && exp_avg && exp_avg_sq && chunks && act_step && group_of && hyper && n_chunks > 0,
                UC2_ERR_ARG, "adamw_step: bad args");
    UC2_REQUIRE(aligned16(param) && aligned16(grad) && aligned16(exp_avg) && aligned16(exp_avg_sq), UC2_ERR_ARG,
                "adamw_step: arenas must be 16-byte aligned");
    UC2_REQUIRE(hyper->global_step >= 1, UC2_ERR_ARG, "adamw_step: global_step starts at 1");
    const int blocks = n_chunks < 16 * num_sms() ? n_chunks : 16 * num_sms();
    adamw_kernel<<<blocks, OPT_THREADS, 0, (cudaStream_t)stream>>>(param, grad, exp_avg, exp_avg_sq, (bf16*)shadow_bf16,
                                                                  chunks, n_chunks, act_step, group_of, *hyper, sqnorm);
    return check_last("adamw_kernel");
}
