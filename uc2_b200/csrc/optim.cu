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

// sum of squares of the gradients of ACTIVE tensors (act_step >= 0) into *out (double)
__global__ void __launch_bounds__(OPT_THREADS)
grad_sqnorm_kernel(const float* __restrict__ grad, const uc2_opt_chunk* __restrict__ chunks, int n_chunks,
                   const int* __restrict__ act_step, double* __restrict__ out) {
    __shared__ float red[OPT_THREADS / 32];
    float acc = 0.f;
    for (int c = blockIdx.x; c < n_chunks; c += gridDim.x) {
        const uc2_opt_chunk ch = chunks[c];
        if (act_step[ch.tensor] < 0) continue;
        const float* g = grad + ch.offset;
        for (int i = threadIdx.x * 4; i < ch.n; i += OPT_THREADS * 4) {
            if (i + 4 <= ch.n) {
                const float4 v = *reinterpret_cast<const float4*>(g + i);
                acc += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
            } else {
                for (int k = i; k < ch.n; ++k) acc += g[k] * g[k];
            }
        }
    }
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int w = 0; w < OPT_THREADS / 32; ++w) s += red[w];
        atomicAdd(out, s);
    }
}

__global__ void __launch_bounds__(OPT_THREADS)
adamw_kernel(float* __restrict__ param, float* __restrict__ grad, float* __restrict__ exp_avg,
             float* __restrict__ exp_avg_sq, bf16* __restrict__ shadow, const uc2_opt_chunk* __restrict__ chunks,
             int n_chunks, const int* __restrict__ act_step, const int* __restrict__ group_of,
             const uc2_adamw_hyper h, const double* __restrict__ sqnorm) {
    // gradient clipping coefficient, torch.nn.utils.clip_grad_norm_: max_norm / (total_norm + 1e-6), capped at 1
    float clip = 1.0f;
    if (h.max_grad_norm > 0.f && sqnorm) {
        const float total = (float)sqrt(*sqnorm);
        clip = fminf(h.max_grad_norm / (total + 1e-6f), 1.0f);
    }
    for (int c = blockIdx.x; c < n_chunks; c += gridDim.x) {
        const uc2_opt_chunk ch = chunks[c];
        const int a = act_step[ch.tensor];
        if (a < 0) continue;                           // p.grad is None -> `continue` (adamw.py:52-53)
        const int grp = group_of[ch.tensor];
        const float lr = h.lr[grp], wd = h.weight_decay[grp];
        const int step = h.global_step - a + 1;        // state['step'] of this tensor
        float step_size = lr;
        if (h.correct_bias) {
            const double bc1 = 1.0 - pow((double)h.beta1, (double)step);
            const double bc2 = 1.0 - pow((double)h.beta2, (double)step);
            step_size = (float)((double)lr * sqrt(bc2) / bc1);
        }
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
            for (int k = 0; k < 4; ++k) {
                const float gr = gv[k] * clip;
                mv[k] = mv[k] * h.beta1 + (1.0f - h.beta1) * gr;
                vv[k] = vv[k] * h.beta2 + (1.0f - h.beta2) * gr * gr;
                const float denom = sqrtf(vv[k]) + h.eps;
                pv[k] = pv[k] - step_size * (mv[k] / denom);
                if (wd > 0.f) pv[k] = pv[k] - lr * wd * pv[k];
            }
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

}  // namespace
}  // namespace uc2

using namespace uc2;

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
