// LayerNorm over rows of 768 (fp32 statistics) forward + backward, and a bf16 column-sum (bias gradients).
// Replaces the apex FusedLayerNorm call sites of model/layer.py:108,149,196,242 and the head transforms
// in model/model.py:1148,1164.  The input row is bf16 (heads) or fp32 (the encoder keeps its residual
// stream -- the LayerNorm inputs and outputs -- in fp32 so that rounding does not accumulate across the
// 12 layers; the bf16 copy of the output is the tensor-core operand of the next GEMM).
// HBM-bound: forward one warp per row, backward two warps per row, 16/32-byte accesses.
#include "common.cuh"
#include "ptx.cuh"

namespace uc2 {
namespace {

constexpr int VPL = 24;
constexpr int LN_WARPS = 8;

__device__ __forceinline__ int col_of(int lane, int i) { return i * 256 + lane * 8; }

template <bool F32>
__device__ __forceinline__ void load_row(const void* base, long long row, int lane, float* v) {
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        if (F32) load8_f32(static_cast<const float*>(base) + row * HID + col_of(lane, i), v + 8 * i);
        else     load8_bf16(static_cast<const bf16*>(base) + row * HID + col_of(lane, i), v + 8 * i);
    }
}

__device__ __forceinline__ void stats(const float* v, float eps, float& mean, float& rstd) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) s += v[i];
    mean = warp_sum(s) * (1.0f / HID);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) { const float d = v[i] - mean; q += d * d; }
    rstd = rsqrtf(warp_sum(q) * (1.0f / HID) + eps);
}

template <bool F32>
__global__ void __launch_bounds__(LN_WARPS * 32)
layernorm_fwd_kernel(const void* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                     float eps, bf16* __restrict__ y, float* __restrict__ y32, long long rows) {
    const int lane = threadIdx.x & 31;
    const long long row = (long long)blockIdx.x * LN_WARPS + (threadIdx.x >> 5);
    griddep_sync();
    if (row >= rows) return;
    float v[VPL];
    load_row<F32>(x, row, lane, v);
    float mean, rstd;
    stats(v, eps, mean, rstd);
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        float g[8], b[8], o[8];
        load8_f32(gamma + col_of(lane, i), g);
        load8_f32(beta + col_of(lane, i), b);
#pragma unroll
        for (int k = 0; k < 8; ++k) o[k] = (v[8 * i + k] - mean) * rstd * g[k] + b[k];
        if (y) store8_bf16(y + row * HID + col_of(lane, i), o);
        if (y32) store8_f32(y32 + row * HID + col_of(lane, i), o);
    }
}

// Backward.  TWO warps share a row (384 columns = 12 values per lane each) and walk rows with a grid stride,
// keeping their halves of the dgamma/dbeta/dbias partial sums in registers (36 accumulators + the half row: under
// 128 registers, so 16 warps fit on an SM; the one-warp-per-row version needed ~150 and was instruction-latency
// bound with 8).  The row statistics and the two projections are combined across the pair through shared memory
// and a 64-thread named barrier.  Memory-level parallelism comes from a per-pair ring of LN_STAGES rows in shared
// memory filled by 1-D bulk copies (cp.async.bulk, completion on an mbarrier per slot): 8 pairs x 4 rows x 4.5 KB
// = 147 KB in flight per SM, which is what it takes to keep HBM busy (the register-only version reached 1.9 TB/s).
constexpr int LN_STAGES = 4;
constexpr int LNB_PAIRS = 8;
constexpr int LNB_WARPS = 2 * LNB_PAIRS;
constexpr int HV = 12;                                 // values per lane of one half row

__device__ __forceinline__ int col_h(int half, int lane, int j) { return half * (HID / 2) + j * 128 + lane * 4; }

__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}

__device__ __forceinline__ void pair_sync(int pair) { asm volatile("bar.sync %0, 64;" ::"r"(pair + 1) : "memory"); }

__device__ __forceinline__ void load4_bf16(const bf16* p, float* f) {
    const uint2 u = *reinterpret_cast<const uint2*>(p);
    const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.x));
    const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.y));
    f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y;
}

__device__ __forceinline__ void store4_bf16(bf16* p, const float* f) {
    uint2 u;
    *reinterpret_cast<__nv_bfloat162*>(&u.x) = __floats2bfloat162_rn(f[0], f[1]);
    *reinterpret_cast<__nv_bfloat162*>(&u.y) = __floats2bfloat162_rn(f[2], f[3]);
    *reinterpret_cast<uint2*>(p) = u;
}

template <bool F32>
__global__ void __launch_bounds__(LNB_WARPS * 32, 1)
layernorm_bwd_kernel(const void* __restrict__ x, const bf16* __restrict__ dy, const float* __restrict__ gamma,
                     float eps, bf16* __restrict__ dx, float* __restrict__ dgamma, float* __restrict__ dbeta,
                     float* __restrict__ dbias, long long rows, bf16* __restrict__ dxm, DropCfg drop) {
    constexpr int XB = F32 ? HID * 4 : HID * 2;       // bytes of one x row
    constexpr int SLOT = XB + HID * 2;                // x row + dy row
    extern __shared__ __align__(128) uint8_t ln_smem[];
    __shared__ float red[3 * HID];
    __shared__ float4 part_a[2][LNB_PAIRS][2];      // double buffered by row parity: one barrier per row is enough
    __shared__ __align__(8) unsigned long long bars[LNB_PAIRS * LN_STAGES];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int pair = warp >> 1, half = warp & 1;
    for (int i = threadIdx.x; i < 3 * HID; i += blockDim.x) red[i] = 0.f;
    const uint32_t ring = ptx::smem_u32(ln_smem) + pair * LN_STAGES * SLOT;
    const uint32_t bar0 = ptx::smem_u32(bars) + pair * LN_STAGES * 8;
    const bool loader = half == 0 && lane == 0;
    if (loader) {
        for (int s = 0; s < LN_STAGES; ++s) ptx::mbar_init(bar0 + 8 * s, 1);
        ptx::fence_barrier_init();
        ptx::fence_proxy_async();
    }
    __syncthreads();
    griddep_sync();
    const long long stride = (long long)gridDim.x * LNB_PAIRS;
    const long long first = (long long)blockIdx.x * LNB_PAIRS + pair;
    auto issue = [&](long long row, int s) {      // loader lane only
        ptx::mbar_arrive_expect_tx(bar0 + 8 * s, SLOT);
        bulk_load(ring + s * SLOT, static_cast<const uint8_t*>(x) + row * XB, XB, bar0 + 8 * s);
        bulk_load(ring + s * SLOT + XB, dy + row * HID, HID * 2, bar0 + 8 * s);
    };
    if (loader)
        for (int s = 0; s < LN_STAGES; ++s)
            if (first + s * stride < rows) issue(first + s * stride, s);
    float g[HV];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        const float4 t = *reinterpret_cast<const float4*>(gamma + col_h(half, lane, j));
        g[4 * j] = t.x; g[4 * j + 1] = t.y; g[4 * j + 2] = t.z; g[4 * j + 3] = t.w;
    }
    float ag[HV], ab[HV], ax[HV];
#pragma unroll
    for (int i = 0; i < HV; ++i) ag[i] = ab[i] = ax[i] = 0.f;
    int s = 0, par = 0;
    uint32_t phase = 0;
    for (long long row = first; row < rows; row += stride, par ^= 1) {
        ptx::mbar_wait(bar0 + 8 * s, phase);
        float v[HV], d[HV];
        const uint8_t* slot = ln_smem + (pair * LN_STAGES + s) * SLOT;
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            const int c = col_h(half, lane, j);
            if (F32) {
                const float4 t = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(slot) + c);
                v[4 * j] = t.x; v[4 * j + 1] = t.y; v[4 * j + 2] = t.z; v[4 * j + 3] = t.w;
            } else {
                load4_bf16(reinterpret_cast<const bf16*>(slot) + c, v + 4 * j);
            }
            load4_bf16(reinterpret_cast<const bf16*>(slot + XB) + c, d + 4 * j);
        }
        // ONE exchange per row: the row statistics (sum x, sum x^2) and both projections of the backward,
        //   s1 = mean(g dy)   and   s2 = mean(g dy xhat) = rstd (sum(g dy x) - mean sum(g dy)) / H,
        // are all sums over the raw row, so the pair needs a single reduction and a single barrier per row (the two-phase
        // form -- statistics, then projections on xhat -- put two dependent shuffle trees and two barriers on every row's
        // latency chain, and this kernel is latency-bound: profiles/r02b_gelu_gemm_lnbwd_ncu.txt)
        float sx = 0.f, sq = 0.f, sg = 0.f, sgx = 0.f;
#pragma unroll
        for (int i = 0; i < HV; ++i) {
            sx += v[i];
            sq += v[i] * v[i];
            const float dg = d[i] * g[i];
            sg += dg;
            sgx += dg * v[i];
        }
        sx = warp_sum(sx);
        sq = warp_sum(sq);
        sg = warp_sum(sg);
        sgx = warp_sum(sgx);
        if (lane == 0) part_a[par][pair][half] = make_float4(sx, sq, sg, sgx);
        pair_sync(pair);                               // also: both warps hold their copy, the slot can be refilled
        if (loader && row + LN_STAGES * stride < rows) {
            ptx::fence_proxy_async();
            issue(row + LN_STAGES * stride, s);
        }
        if (++s == LN_STAGES) { s = 0; phase ^= 1u; }
        const float4 pa0 = part_a[par][pair][0], pa1 = part_a[par][pair][1];
        const float mean = (pa0.x + pa1.x) * (1.0f / HID);
        const float var = fmaxf((pa0.y + pa1.y) * (1.0f / HID) - mean * mean, 0.f);
        const float rstd = rsqrtf(var + eps);
        const float sum_g = pa0.z + pa1.z;
        const float s1 = sum_g * (1.0f / HID);
        const float s2 = rstd * ((pa0.w + pa1.w) - mean * sum_g) * (1.0f / HID);
#pragma unroll
        for (int i = 0; i < HV; ++i) {
            v[i] = (v[i] - mean) * rstd;          // xhat
            ag[i] += d[i] * v[i];
            ab[i] += d[i];
            d[i] = rstd * (d[i] * g[i] - s1 - v[i] * s2);
        }
#pragma unroll
        for (int j = 0; j < 3; ++j) store4_bf16(dx + row * HID + col_h(half, lane, j), d + 4 * j);
        if (dxm) {
            // x was dropout(dense) + residual: the dense branch (its wgrad / dgrad / bias gradient) sees the masked,
            // rescaled gradient, the residual branch the plain one written above
            const uint32_t i0 = static_cast<uint32_t>(row) * HID;
#pragma unroll
            for (int i = 0; i < HV; i += 2) {
                bool k0, k1;
                drop_keep2(drop.key, i0 + col_h(half, lane, i >> 2) + (i & 3), drop.thresh, k0, k1);
                d[i] = k0 ? d[i] * drop.scale : 0.f;
                d[i + 1] = k1 ? d[i + 1] * drop.scale : 0.f;
            }
#pragma unroll
            for (int j = 0; j < 3; ++j) store4_bf16(dxm + row * HID + col_h(half, lane, j), d + 4 * j);
        }
#pragma unroll
        for (int i = 0; i < HV; ++i) ax[i] += d[i];
    }
#pragma unroll
    for (int i = 0; i < HV; ++i) {
        const int c = col_h(half, lane, i >> 2) + (i & 3);
        atomicAdd(red + c, ag[i]);
        atomicAdd(red + HID + c, ab[i]);
        atomicAdd(red + 2 * HID + c, ax[i]);
    }
    __syncthreads();
    for (int c = threadIdx.x; c < HID; c += blockDim.x) {
        atomicAdd(dgamma + c, red[c]);
        atomicAdd(dbeta + c, red[HID + c]);
        if (dbias) atomicAdd(dbias + c, red[2 * HID + c]);
    }
}

// column sums of a bf16 matrix: CTA = 8 row phases x 256 columns (8 per thread)
__global__ void __launch_bounds__(256)
colsum_bf16_kernel(const bf16* __restrict__ x, long long ld, long long rows, int cols, float* __restrict__ out,
                   int row_chunk) {
    griddep_sync();
    const int cg = blockIdx.x * 256 + (threadIdx.x & 31) * 8;      // 8 consecutive columns per thread
    const int rsub = threadIdx.x >> 5;                              // 8 row phases
    const long long r0 = (long long)blockIdx.y * row_chunk;
    const long long r1 = r0 + row_chunk < rows ? r0 + row_chunk : rows;
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (cg + 8 <= cols) {
        for (long long r = r0 + rsub; r < r1; r += 8) {
            float v[8];
            load8_bf16(x + r * ld + cg, v);
#pragma unroll
            for (int k = 0; k < 8; ++k) acc[k] += v[k];
        }
    } else if (cg < cols) {
        for (long long r = r0 + rsub; r < r1; r += 8)
            for (int k = 0; k < 8 && cg + k < cols; ++k) acc[k] += __bfloat162float(x[r * ld + cg + k]);
    }
    __shared__ float red[8][256 + 8];
#pragma unroll
    for (int k = 0; k < 8; ++k) red[rsub][(threadIdx.x & 31) * 8 + k] = acc[k];
    __syncthreads();
    const int c = threadIdx.x;
    float s = 0.f;
#pragma unroll
    for (int r = 0; r < 8; ++r) s += red[r][c];
    const int gc = blockIdx.x * 256 + c;
    if (gc < cols && s != 0.f) atomicAdd(out + gc, s);
}

}  // namespace
}  // namespace uc2

using namespace uc2;

extern "C" UC2_API int uc2_layernorm_fwd(const void* x, int x_is_f32, const float* gamma, const float* beta, float eps,
                                         void* y_bf16, float* y_f32, long long rows, void* stream) {
    if (int rc = require_sm100()) return rc;
    UC2_REQUIRE(x && gamma && beta && (y_bf16 || y_f32) && rows > 0, UC2_ERR_ARG, "layernorm_fwd: bad args");
    UC2_REQUIRE(aligned16(x) && aligned16(y_bf16) && aligned16(y_f32) && aligned16(gamma) && aligned16(beta), UC2_ERR_ARG,
                "layernorm_fwd: pointers must be 16-byte aligned");
    const unsigned blocks = (unsigned)((rows + LN_WARPS - 1) / LN_WARPS);
    if (x_is_f32)
        launch_pdl(layernorm_fwd_kernel<true>, dim3(blocks), dim3(LN_WARPS * 32), 0, (cudaStream_t)stream, 1, x, gamma,
                   beta, eps, (bf16*)y_bf16, y_f32, rows);
    else
        launch_pdl(layernorm_fwd_kernel<false>, dim3(blocks), dim3(LN_WARPS * 32), 0, (cudaStream_t)stream, 1, x, gamma,
                   beta, eps, (bf16*)y_bf16, y_f32, rows);
    return check_last("layernorm_fwd_kernel");
}

extern "C" UC2_API int uc2_layernorm_bwd(const void* x, int x_is_f32, const void* dy, const float* gamma, float eps,
                                         void* dx, float* dgamma, float* dbeta, float* dbias, long long rows,
                                         void* stream) {
    return uc2_layernorm_bwd_dropout(x, x_is_f32, dy, gamma, eps, dx, dgamma, dbeta, dbias, rows, nullptr, 0u, 0u, 1.f,
                                     stream);
}

extern "C" UC2_API int uc2_layernorm_bwd_dropout(const void* x, int x_is_f32, const void* dy, const float* gamma,
                                                 float eps, void* dx, float* dgamma, float* dbeta, float* dbias,
                                                 long long rows, void* dx_masked, unsigned int drop_key,
                                                 unsigned int drop_thresh, float drop_scale, void* stream) {
    if (int rc = require_sm100()) return rc;
    UC2_REQUIRE((dx_masked != nullptr) == (drop_thresh != 0) && drop_thresh < 65536u && aligned16(dx_masked), UC2_ERR_ARG,
                "layernorm_bwd: dx_masked goes with drop_thresh in (0, 65536)");
    const DropCfg drop = {drop_key, drop_thresh, drop_scale};
    UC2_REQUIRE(x && dy && gamma && dx && dgamma && dbeta && rows > 0, UC2_ERR_ARG, "layernorm_bwd: bad args");
    UC2_REQUIRE(aligned16(x) && aligned16(dy) && aligned16(dx) && aligned16(gamma), UC2_ERR_ARG,
                "layernorm_bwd: pointers must be 16-byte aligned");
    long long blocks = (rows + LNB_PAIRS - 1) / LNB_PAIRS;
    const long long cap = num_sms();                  // one CTA per SM: the row ring takes most of its shared memory
    if (blocks > cap) blocks = cap;
    const int smem = LNB_PAIRS * LN_STAGES * ((x_is_f32 ? HID * 4 : HID * 2) + HID * 2);
    if (x_is_f32) {
        UC2_CUDA(cudaFuncSetAttribute(layernorm_bwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        launch_pdl(layernorm_bwd_kernel<true>, dim3((unsigned)blocks), dim3(LNB_WARPS * 32), smem, (cudaStream_t)stream, 1,
                   x, (const bf16*)dy, gamma, eps, (bf16*)dx, dgamma, dbeta, dbias, rows, (bf16*)dx_masked, drop);
    } else {
        UC2_CUDA(cudaFuncSetAttribute(layernorm_bwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        launch_pdl(layernorm_bwd_kernel<false>, dim3((unsigned)blocks), dim3(LNB_WARPS * 32), smem, (cudaStream_t)stream, 1,
                   x, (const bf16*)dy, gamma, eps, (bf16*)dx, dgamma, dbeta, dbias, rows, (bf16*)dx_masked, drop);
    }
    return check_last("layernorm_bwd_kernel");
}

extern "C" UC2_API int uc2_colsum_bf16(const void* x, long long ld, long long rows, int cols, float* out,
                                       void* stream) {
    if (int rc = require_sm100()) return rc;
    UC2_REQUIRE(x && out && rows > 0 && cols > 0 && ld >= cols, UC2_ERR_ARG, "colsum_bf16: bad args");
    UC2_REQUIRE((ld % 8 == 0) && aligned16(x), UC2_ERR_ARG, "colsum_bf16: x must be 16-byte aligned with ld %% 8 == 0");
    const int col_blocks = (cols + 255) / 256;
    int row_blocks = (int)((4LL * num_sms() + col_blocks - 1) / col_blocks);
    long long row_chunk = (rows + row_blocks - 1) / row_blocks;
    if (row_chunk < 64) row_chunk = 64;
    row_blocks = (int)((rows + row_chunk - 1) / row_chunk);
    launch_pdl(colsum_bf16_kernel, dim3(col_blocks, row_blocks), dim3(256), 0, (cudaStream_t)stream, 1, (const bf16*)x,
               ld, rows, cols, out, (int)row_chunk);
    return check_last("colsum_bf16_kernel");
}
