// Shared device/host helpers for the uc2_b200 sm_100a kernels.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/uc2_b200.h"

namespace uc2 {

// ---------------------------------------------------------------------------------------------
// error plumbing (never throw across the C ABI)
// ---------------------------------------------------------------------------------------------
void set_error(const char* fmt, ...);
int check_last(const char* what);   // cudaGetLastError -> UC2_ERR_CUDA

#define UC2_REQUIRE(cond, code, ...)            \
    do {                                        \
        if (!(cond)) {                          \
            uc2::set_error(__VA_ARGS__);        \
            return (code);                      \
        }                                       \
    } while (0)

#define UC2_CUDA(expr)                                                              \
    do {                                                                            \
        cudaError_t _e = (expr);                                                    \
        if (_e != cudaSuccess) {                                                    \
            uc2::set_error("%s failed: %s", #expr, cudaGetErrorString(_e));         \
            return UC2_ERR_CUDA;                                                    \
        }                                                                           \
    } while (0)

int num_sms();

// RAII launch timer (no-op unless uc2_profile_enable(1)); kind 0 = GEMM, 1 = attention, 2 = other
struct ProfScope {
    ProfScope(cudaStream_t s, int kind, double work);
    ~ProfScope();
    cudaStream_t stream_;
    int idx_;
};
int require_sm100();
bool pdl_enabled();      // programmatic dependent launch (off with UC2_NO_PDL=1)
bool gemm_sched_dynamic();   // persistent GEMM workers draw tiles from a counter (uc2_gemm_sched_dynamic)
bool attn_tc_enabled();  // tcgen05 attention kernels (attention_tc.cu) in use; UC2_ATTN_TCGEN05=0 switches them off
inline bool attn_tc_bwd_serves(int S) { return (S + 15) / 16 * 16 <= 256; }   // shapes uc2_attention_bwd_tc takes

// Launch `kern` allowing it to start while the previous kernel of the stream is still draining (programmatic
// dependent launch): its CTAs may become resident and run their prologue early, and MUST call griddep_wait()
// before their first global access that depends on (or overwrites data of) the previous kernel.
template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                                     int cluster_x, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute at[2];
    int n = 0;
    if (cluster_x > 1) {
        at[n].id = cudaLaunchAttributeClusterDimension;
        at[n].val.clusterDim.x = cluster_x; at[n].val.clusterDim.y = 1; at[n].val.clusterDim.z = 1;
        ++n;
    }
    if (pdl_enabled()) {
        at[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[n].val.programmaticStreamSerializationAllowed = 1;
        ++n;
    }
    cfg.attrs = at;
    cfg.numAttrs = n;
    return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

// cuTensorMapEncodeTiled through the runtime's driver entry point lookup (no link-time libcuda dependency)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn tensormap_encoder();
// 2-D bf16 tensor [rows][cols], row pitch ld elements; box = [box_rows][64 columns], SWIZZLE_128B
int make_tmap(CUtensorMap* m, const void* base, long long rows, long long cols, long long ld, int box_rows);

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// ---------------------------------------------------------------------------------------------
// small device utilities
// ---------------------------------------------------------------------------------------------
#ifdef __CUDACC__

typedef __nv_bfloat16 bf16;
typedef __nv_bfloat162 bf162;

// Programmatic dependent launch, device side: wait for the previous kernel of the stream (completion + memory
// visibility), then let the next kernel's CTAs be scheduled as soon as every CTA of this grid got here or exited.
__device__ __forceinline__ void griddep_sync() {
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

constexpr int HID = 768;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
    bf162 v = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float2 unpack_bf16(uint32_t u) {
    bf162 v = *reinterpret_cast<bf162*>(&u);
    return __bfloat1622float2(v);
}

// 8 bf16 <-> 8 floats through one 16-byte access
__device__ __forceinline__ void load8_bf16(const bf16* p, float* f) {
    uint4 u = *reinterpret_cast<const uint4*>(p);
    float2 a = unpack_bf16(u.x), b = unpack_bf16(u.y), c = unpack_bf16(u.z), d = unpack_bf16(u.w);
    f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
}
__device__ __forceinline__ void store8_bf16(bf16* p, const float* f) {
    uint4 u;
    u.x = pack_bf16(f[0], f[1]); u.y = pack_bf16(f[2], f[3]);
    u.z = pack_bf16(f[4], f[5]); u.w = pack_bf16(f[6], f[7]);
    *reinterpret_cast<uint4*>(p) = u;
}
__device__ __forceinline__ void load8_f32(const float* p, float* f) {
    float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
    f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}
__device__ __forceinline__ void store8_f32(float* p, const float* f) {
    *reinterpret_cast<float4*>(p) = make_float4(f[0], f[1], f[2], f[3]);
    *reinterpret_cast<float4*>(p + 4) = make_float4(f[4], f[5], f[6], f[7]);
}

// Counter-based dropout (host mirror and rationale: uc2_b200/dropout.py).  One 32-bit mix serves TWO neighbouring
// elements: keep(idx) <=> the (idx & 1)-th 16-bit half of lowbias32((idx >> 1) ^ site key) is >= thresh =
// round(p * 65536); forward and backward regenerate the same mask.  Kernels that hold element pairs (2j, 2j + 1) --
// every site does, along its fastest index -- pay one mix per pair (drop_keep2).
struct DropCfg {
    uint32_t key;
    uint32_t thresh;     // 0: dropout off
    float scale;         // 1 / (1 - p)
};
__host__ __device__ __forceinline__ uint32_t lowbias32(uint32_t x) {
    x ^= x >> 16; x *= 0x7FEB352Du;
    x ^= x >> 15; x *= 0x846CA68Bu;
    x ^= x >> 16;
    return x;
}
__device__ __forceinline__ bool drop_keep(uint32_t key, uint32_t idx, uint32_t thresh) {
    const uint32_t h = lowbias32((idx >> 1) ^ key);
    return ((idx & 1u) ? (h >> 16) : (h & 0xFFFFu)) >= thresh;
}
// keep flags of elements idx and idx + 1; one mix when idx is even (the common, aligned case)
__device__ __forceinline__ void drop_keep2(uint32_t key, uint32_t idx, uint32_t thresh, bool& k0, bool& k1) {
    if ((idx & 1u) == 0u) {
        const uint32_t h = lowbias32((idx >> 1) ^ key);
        k0 = (h & 0xFFFFu) >= thresh;
        k1 = (h >> 16) >= thresh;
    } else {
        k0 = drop_keep(key, idx, thresh);
        k1 = drop_keep(key, idx + 1u, thresh);
    }
}
__host__ __device__ __forceinline__ uint32_t drop_head_key(uint32_t key, uint32_t bh) {
    return lowbias32(key ^ (bh * 0x9E3779B9u + 0x7F4A7C15u));
}

// Attention-probability dropout (layer.py:94) of the tcgen05 kernels.  The forward walks a query ROW (thread = query,
// loop over keys), the backward a key ROW (thread = key, loop over queries); a per-element mix would cost more than the
// softmax arithmetic itself in both.  So the [S x S] mask of a (batch, head) is tiled in 16 x 16 blocks: one lowbias32
// per block, then element (i, j) takes  e = h_block * CA^(i & 15) * CB^(j & 15)  (mod 2^32; CA, CB odd, so every
// element's e is a bijection of the block hash and exactly uniform) and is kept iff e >= thresh << 16.  Whichever
// index a thread owns contributes a per-thread factor, the other one compile-time immediates: one IMAD + one ISETP
// per element in either orientation.  Host mirror: uc2_b200/dropout.py attn_keep_mask_np.
constexpr uint32_t DROP_CA = 0x9E3779B1u, DROP_CB = 0x85EBCA77u;
__host__ __device__ constexpr uint32_t drop_pow(uint32_t c, int k) {
    uint32_t r = 1u;
    for (int i = 0; i < k; ++i) r *= c;
    return r;
}
// c^(k & 15) for a run-time k
__device__ __forceinline__ uint32_t drop_pow_rt(uint32_t c, uint32_t k) {
    const uint32_t c2 = c * c, c4 = c2 * c2, c8 = c4 * c4;
    uint32_t r = (k & 1u) ? c : 1u;
    r *= (k & 2u) ? c2 : 1u;
    r *= (k & 4u) ? c4 : 1u;
    r *= (k & 8u) ? c8 : 1u;
    return r;
}
__device__ __forceinline__ uint32_t drop_block_hash(uint32_t hkey, uint32_t iblk, uint32_t jblk) {
    return lowbias32(hkey ^ ((iblk << 16) | jblk));
}

// erf-form GELU (model/layer.py:31-37) and its derivative.
// 0.5 erfc(|x|/sqrt2) through Abramowitz-Stegun 7.1.25: erfc(z) = t (a1 + t (a2 + t a3)) exp(-z^2),
// t = 1 / (1 + p z), z >= 0, |error| <= 2.5e-5 on erf.  The results are rounded to bf16 (half an ulp is 2e-3
// relative), so gelu is off by at most 2.6e-5 absolute / gelu' by 1.1e-5 -- two orders below the storage
// rounding -- while the whole evaluation is 11 FP32 instructions + MUFU.RCP + MUFU.EX2 (erff() costs ~3x that,
// and these run in GEMM epilogues that have to keep up with the tensor pipe).  Written on the erfc side:
// gelu(x) = max(x, 0) - |x| h has no 1 + erf cancellation in the negative tail.  exp(-z^2) = exp(-x^2/2) is
// shared with the density term of the derivative.
__device__ __forceinline__ float fast_rcp(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float fast_ex2(float x) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
// tanh(x) = 1 - 2 / (exp(2x) + 1): two MUFU ops, absolute error ~1e-6 (BertPooler, model/layer.py:184)
__device__ __forceinline__ float fast_tanh(float x) {
    const float e = fast_ex2(fminf(2.8853900817779268f * x, 80.f));
    return 1.0f - 2.0f * fast_rcp(e + 1.0f);
}
struct GeluTerms { float h, e; };       // h = 0.5 erfc(|x|/sqrt2), e = exp(-x^2/2)
__device__ __forceinline__ GeluTerms gelu_terms(float x) {
    const float t = fast_rcp(fmaf(0.47047f * 0.70710678118654752f, fabsf(x), 1.0f));
    float pl = fmaf(0.5f * 0.7478556f, t, 0.5f * -0.0958798f);
    pl = fmaf(pl, t, 0.5f * 0.3480242f);
    GeluTerms g;
    g.e = fast_ex2((x * x) * (-0.5f * 1.4426950408889634f));
    g.h = (pl * t) * g.e;
    return g;
}
__device__ __forceinline__ float gelu_erf(float x) { return fmaf(-fabsf(x), gelu_terms(x).h, fmaxf(x, 0.f)); }
__device__ __forceinline__ float gelu_erf_grad(float x) {
    const GeluTerms g = gelu_terms(x);
    const float Phi = 0.5f + copysignf(0.5f - g.h, x);
    return fmaf(x * 0.3989422804014327f, g.e, Phi);
}

// ---- packed fp32 pairs: Blackwell executes fma / mul / add on two fp32 values per lane in one instruction (FFMA2,
// FMUL2, FADD2; each half is rounded exactly like the scalar instruction).  The GELU epilogues of the GEMM are bound
// by the FP32 pipe, not by the tensor pipe they are meant to hide under; in pairs they cost half the pipe time.
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk2(float a, float b) {
    f32x2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ f32x2 pk2u(uint32_t a, uint32_t b) {
    f32x2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(a), "r"(b));
    return r;
}
__device__ __forceinline__ void up2(f32x2 v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
    f32x2 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) {
    f32x2 r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) {
    f32x2 r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f32x2 splat2(float a) { return pk2(a, a); }
__device__ __forceinline__ uint32_t pack_bf16x2(f32x2 v) {
    float a, b;
    up2(v, a, b);
    return pack_bf16(a, b);
}

// gelu_terms() on a pair, the same operations in the same order (bit-identical halves).  nax = -|x|: its product
// with -k is the scalar form's k |x|, and it is the multiplier of gelu's last fma.
struct GeluTerms2 { f32x2 h, e, nax; };
__device__ __forceinline__ GeluTerms2 gelu_terms2(f32x2 x) {
    float x0, x1;
    up2(x, x0, x1);
    GeluTerms2 g;
    g.nax = pk2(-fabsf(x0), -fabsf(x1));
    const f32x2 d = fma2(g.nax, splat2(-(0.47047f * 0.70710678118654752f)), splat2(1.0f));
    float d0, d1;
    up2(d, d0, d1);
    const f32x2 t = pk2(fast_rcp(d0), fast_rcp(d1));
    f32x2 pl = fma2(splat2(0.5f * 0.7478556f), t, splat2(0.5f * -0.0958798f));
    pl = fma2(pl, t, splat2(0.5f * 0.3480242f));
    const f32x2 xx = mul2(mul2(x, x), splat2(-0.5f * 1.4426950408889634f));
    float q0, q1;
    up2(xx, q0, q1);
    g.e = pk2(fast_ex2(q0), fast_ex2(q1));
    g.h = mul2(mul2(pl, t), g.e);
    return g;
}
__device__ __forceinline__ f32x2 gelu_erf2(f32x2 x) {
    const GeluTerms2 g = gelu_terms2(x);
    float x0, x1;
    up2(x, x0, x1);
    return fma2(g.nax, g.h, pk2(fmaxf(x0, 0.f), fmaxf(x1, 0.f)));
}
__device__ __forceinline__ f32x2 gelu_erf_grad2(f32x2 x) {
    const GeluTerms2 g = gelu_terms2(x);
    const f32x2 s = fma2(g.h, splat2(-1.0f), splat2(0.5f));          // 0.5 - h, one rounding like the scalar form
    float x0, x1, s0, s1;
    up2(x, x0, x1);
    up2(s, s0, s1);
    const f32x2 Phi = add2(splat2(0.5f), pk2(copysignf(s0, x0), copysignf(s1, x1)));
    return fma2(mul2(x, splat2(0.3989422804014327f)), g.e, Phi);
}

#endif  // __CUDACC__

}  // namespace uc2
