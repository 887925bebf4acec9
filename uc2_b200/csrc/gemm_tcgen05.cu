// Persistent, warp-specialised bf16 GEMM for sm_100a:
//   TMA (cp.async.bulk.tensor, SWIZZLE_128B) -> shared memory ring -> tcgen05.mma (fp32 accumulators
//   in TMEM, double buffered) -> tcgen05.ld epilogue with fused bias / GELU / dGELU / tanh / residual.
// One kernel template covers forward (A,B K-major), dgrad (B MN-major) and wgrad (A,B MN-major,
// split-K with fp32 atomic accumulation).  Replaces the nn.Linear call sites listed in
// include/uc2_b200.h (model/layer.py:76-78,112,140,153; model/model.py:359,1153-1169).
//
// Warp roles (320 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer,
// warps 2..9 = epilogue.  An epilogue warp owns TMEM lane quarter warp_idx % 4 (the hardware's rule) and
// every other 32-column chunk of the accumulator; a thread owns one output row of a chunk, so all the
// residual / aux / bias loads of a chunk are in flight together (no shared-memory transpose) while the
// tcgen05.ld of the accumulator is outstanding, and the epilogue of tile i hides under the MMAs of tile i+1.
#include <mutex>

#include "common.cuh"
#include "ptx.cuh"

namespace uc2 {

namespace {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;    // 64 bf16 = 128 B = one SWIZZLE_128B row
constexpr int UMMA_K = 16;
constexpr int EPI_WARPS = 8;
constexpr int GEMM_THREADS = 64 + 32 * EPI_WARPS;
constexpr int EPI_COLS = 32;   // accumulator columns per tcgen05.ld (one fp32 row slice of 128 B per thread)
constexpr int STAGING_BYTES = 0;
constexpr int SMEM_LIMIT = 232448;   // 227 KB opt-in maximum per CTA

template <int BLOCK_N>
struct Cfg {
    static constexpr int A_BYTES = BLOCK_M * BLOCK_K * 2;
    static constexpr int B_BYTES = BLOCK_N * BLOCK_K * 2;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int MAX_STAGES = (SMEM_LIMIT - 1024 - STAGING_BYTES - 256) / STAGE_BYTES;
    static constexpr int STAGES = MAX_STAGES > 8 ? 8 : MAX_STAGES;
    static constexpr int TMEM_COLS = 2 * BLOCK_N;   // two accumulator buffers; 128/256/512: power of two
    static constexpr int SMEM_BYTES = 1024 + STAGES * STAGE_BYTES + STAGING_BYTES + 256;
};

struct GemmParams {
    int M, N, K;
    int num_m_blocks, num_n_blocks, split_k, k_blocks_per_split, num_k_blocks;
    const float* bias;
    const bf16* residual; long long ld_res; int res_f32;
    const bf16* aux; long long ld_aux;
    int act;
    bf16* out_bf16; long long ld_out;
    bf16* out_pre; long long ld_pre;
    float* out_f32; long long ld_f32;
    int accumulate;
    int vec_ok;   // leading dimensions / pointers allow 16-byte row-slice access -> vector epilogue
};

__device__ __forceinline__ float apply_act(float v, int act, float aux) {
    if (act == UC2_ACT_GELU) return gelu_erf(v);
    if (act == UC2_ACT_DGELU) return v * gelu_erf_grad(aux);
    if (act == UC2_ACT_TANH) return tanhf(v);
    return v;
}

template <int BLOCK_N, bool A_MN, bool B_MN>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                 const GemmParams p) {
    using C = Cfg<BLOCK_N>;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem_gen = smem_raw + (smem_base - ptx::smem_u32(smem_raw));
    const uint32_t bar_base = smem_base + C::STAGES * C::STAGE_BYTES + STAGING_BYTES;
    // barrier layout (8 B each): full[STAGES], empty[STAGES], tmem_full[2], tmem_empty[2], then tmem ptr
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (C::STAGES + s); };
    auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * C::STAGES + a); };
    auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * C::STAGES + 2 + a); };
    const uint32_t tmem_ptr_addr = bar_base + 8u * (2 * C::STAGES + 4);
    volatile uint32_t* tmem_ptr_gen =
        reinterpret_cast<volatile uint32_t*>(smem_gen + C::STAGES * C::STAGE_BYTES + STAGING_BYTES +
                                             8 * (2 * C::STAGES + 4));

    const int warp_idx = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        ptx::prefetch_tmap(&tmap_a);
        ptx::prefetch_tmap(&tmap_b);
        for (int s = 0; s < C::STAGES; ++s) {
            ptx::mbar_init(full_bar(s), 1);
            ptx::mbar_init(empty_bar(s), 1);
        }
        for (int a = 0; a < 2; ++a) {
            ptx::mbar_init(tfull_bar(a), 1);
            ptx::mbar_init(tempty_bar(a), EPI_WARPS);   // one arrive per epilogue warp
        }
        ptx::fence_barrier_init();
        ptx::fence_proxy_async();
    }
    if (warp_idx == 1) {
        ptx::tmem_alloc(tmem_ptr_addr, C::TMEM_COLS);
        ptx::tmem_relinquish();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_gen;

    const int tiles_mn = p.num_m_blocks * p.num_n_blocks;
    const int total_tiles = tiles_mn * p.split_k;

    if (warp_idx == 0) {
        // ===================================== TMA producer =====================================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
                const int n_blk = t % p.num_n_blocks;
                const int m_blk = (t / p.num_n_blocks) % p.num_m_blocks;
                const int split = t / tiles_mn;
                const int kb0 = split * p.k_blocks_per_split;
                const int kb1 = min(kb0 + p.k_blocks_per_split, p.num_k_blocks);
                for (int kb = kb0; kb < kb1; ++kb) {
                    ptx::mbar_wait(empty_bar(stage), phase ^ 1u);
                    const uint32_t sa = smem_base + stage * C::STAGE_BYTES;
                    const uint32_t sb = sa + C::A_BYTES;
                    ptx::mbar_arrive_expect_tx(full_bar(stage), C::STAGE_BYTES);
                    if (!A_MN) {
                        ptx::tma_load_2d(sa, &tmap_a, full_bar(stage), kb * BLOCK_K, m_blk * BLOCK_M);
                    } else {
#pragma unroll
                        for (int j = 0; j < BLOCK_M / 64; ++j)
                            ptx::tma_load_2d(sa + j * 8192, &tmap_a, full_bar(stage), m_blk * BLOCK_M + j * 64,
                                             kb * BLOCK_K);
                    }
                    if (!B_MN) {
                        ptx::tma_load_2d(sb, &tmap_b, full_bar(stage), kb * BLOCK_K, n_blk * BLOCK_N);
                    } else {
#pragma unroll
                        for (int j = 0; j < BLOCK_N / 64; ++j)
                            ptx::tma_load_2d(sb + j * 8192, &tmap_b, full_bar(stage), n_blk * BLOCK_N + j * 64,
                                             kb * BLOCK_K);
                    }
                    if (++stage == C::STAGES) { stage = 0; phase ^= 1u; }
                }
            }
        }
    } else if (warp_idx == 1) {
        // ===================================== MMA issuer =======================================
        if (lane == 0) {
            constexpr uint32_t idesc = ptx::idesc_bf16_f32(BLOCK_M, BLOCK_N, A_MN, B_MN);
            // K-major: 8-row groups 1024 B apart (SBO); MN-major: 64-element blocks 8192 B apart (LBO),
            // 8-k-row groups 1024 B apart (SBO)
            constexpr uint32_t A_LBO = A_MN ? 8192u : 16u, B_LBO = B_MN ? 8192u : 16u;
            constexpr uint32_t A_KSTEP = A_MN ? 2048u : 32u, B_KSTEP = B_MN ? 2048u : 32u;
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
                const int split = t / tiles_mn;
                const int kb0 = split * p.k_blocks_per_split;
                const int kb1 = min(kb0 + p.k_blocks_per_split, p.num_k_blocks);
                ptx::mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
                ptx::tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
                for (int kb = kb0; kb < kb1; ++kb) {
                    ptx::mbar_wait(full_bar(stage), phase);
                    ptx::tc_fence_after();
                    const uint32_t sa = smem_base + stage * C::STAGE_BYTES;
                    const uint32_t sb = sa + C::A_BYTES;
#pragma unroll
                    for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
                        const uint64_t da = ptx::smem_desc_sw128(sa + k * A_KSTEP, A_LBO, 1024u);
                        const uint64_t db = ptx::smem_desc_sw128(sb + k * B_KSTEP, B_LBO, 1024u);
                        ptx::umma_bf16(d_tmem, da, db, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
                    }
                    ptx::umma_commit(empty_bar(stage));   // frees the smem slot once these MMAs retire
                    if (++stage == C::STAGES) { stage = 0; phase ^= 1u; }
                }
                ptx::umma_commit(tfull_bar(acc));          // accumulator complete -> epilogue
                if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
            }
        }
    } else {
        // ===================================== epilogue =========================================
        const int q = warp_idx & 3;                       // TMEM lane quarter owned by this warp
        const int half = (warp_idx - 2) >> 2;             // which of the interleaved 32-column chunks
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
            const int n_blk = t % p.num_n_blocks;
            const int m_blk = (t / p.num_n_blocks) % p.num_m_blocks;
            ptx::mbar_wait(tfull_bar(acc), acc_phase);
            ptx::tc_fence_after();
            const long long grow = (long long)m_blk * BLOCK_M + q * 32 + lane;
            const bool row_ok = grow < p.M;
            constexpr int MY_CHUNKS = BLOCK_N / EPI_COLS / 2;
#pragma unroll 1
            for (int i = 0; i < MY_CHUNKS; ++i) {
                const int c = 2 * i + half;
                const int col0 = n_blk * BLOCK_N + c * EPI_COLS;
                uint32_t r[32];
                ptx::tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BLOCK_N + c * EPI_COLS,
                                   r);
                // operands of the fused epilogue, requested while the accumulator load is in flight
                const bool vec = p.vec_ok && col0 + EPI_COLS <= p.N;       // warp-uniform
                const bool live = row_ok && col0 < p.N;
                uint4 ex[8];                                               // residual (fp32: 8 x 16 B, bf16: 4) or aux
                if (vec && live) {
                    if (p.residual) {
                        if (p.res_f32) {
                            const uint4* src = reinterpret_cast<const uint4*>(
                                reinterpret_cast<const float*>(p.residual) + grow * p.ld_res + col0);
#pragma unroll
                            for (int j = 0; j < 8; ++j) ex[j] = __ldg(src + j);
                        } else {
                            const uint4* src = reinterpret_cast<const uint4*>(p.residual + grow * p.ld_res + col0);
#pragma unroll
                            for (int j = 0; j < 4; ++j) ex[j] = __ldg(src + j);
                        }
                    } else if (p.act == UC2_ACT_DGELU) {
                        const uint4* src = reinterpret_cast<const uint4*>(p.aux + grow * p.ld_aux + col0);
#pragma unroll
                        for (int j = 0; j < 4; ++j) ex[j] = __ldg(src + j);
                    }
                }
                ptx::tmem_wait_ld();
                if (i == MY_CHUNKS - 1) {
                    // all TMEM reads of this accumulator are done: hand it back to the MMA warp
                    ptx::tc_fence_before();
                    __syncwarp();
                    if (lane == 0) ptx::mbar_arrive(tempty_bar(acc));
                }
                if (!live) continue;
                if (vec) {
                    float v[32];
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
                    if (p.bias) {
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + col0) + j);
                            v[4 * j] += b.x; v[4 * j + 1] += b.y; v[4 * j + 2] += b.z; v[4 * j + 3] += b.w;
                        }
                    }
                    if (p.out_pre) {
                        uint4* dst = reinterpret_cast<uint4*>(p.out_pre + grow * p.ld_pre + col0);
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            dst[j] = make_uint4(pack_bf16(v[8 * j], v[8 * j + 1]), pack_bf16(v[8 * j + 2], v[8 * j + 3]),
                                                pack_bf16(v[8 * j + 4], v[8 * j + 5]), pack_bf16(v[8 * j + 6], v[8 * j + 7]));
                    }
                    if (p.act == UC2_ACT_GELU) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = gelu_erf(v[j]);
                    } else if (p.act == UC2_ACT_TANH) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = tanhf(v[j]);
                    } else if (p.act == UC2_ACT_DGELU && !p.residual) {
                        const uint32_t* a = reinterpret_cast<const uint32_t*>(ex);
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            const float2 u = unpack_bf16(a[j]);
                            v[2 * j] *= gelu_erf_grad(u.x);
                            v[2 * j + 1] *= gelu_erf_grad(u.y);
                        }
                    } else if (p.act == UC2_ACT_DGELU) {
                        // residual and aux together (not used by the encoder): aux fetched late
                        const uint32_t* a = reinterpret_cast<const uint32_t*>(p.aux + grow * p.ld_aux + col0);
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            const float2 u = unpack_bf16(__ldg(a + j));
                            v[2 * j] *= gelu_erf_grad(u.x);
                            v[2 * j + 1] *= gelu_erf_grad(u.y);
                        }
                    }
                    if (p.residual) {
                        if (p.res_f32) {
                            const float* f = reinterpret_cast<const float*>(ex);
#pragma unroll
                            for (int j = 0; j < 32; ++j) v[j] += f[j];
                        } else {
                            const uint32_t* a = reinterpret_cast<const uint32_t*>(ex);
#pragma unroll
                            for (int j = 0; j < 16; ++j) {
                                const float2 u = unpack_bf16(a[j]);
                                v[2 * j] += u.x;
                                v[2 * j + 1] += u.y;
                            }
                        }
                    }
                    if (p.out_bf16) {
                        uint4* dst = reinterpret_cast<uint4*>(p.out_bf16 + grow * p.ld_out + col0);
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            dst[j] = make_uint4(pack_bf16(v[8 * j], v[8 * j + 1]), pack_bf16(v[8 * j + 2], v[8 * j + 3]),
                                                pack_bf16(v[8 * j + 4], v[8 * j + 5]), pack_bf16(v[8 * j + 6], v[8 * j + 7]));
                    }
                    if (p.out_f32) {
                        float* dst = p.out_f32 + grow * p.ld_f32 + col0;
                        if (p.accumulate) {
#pragma unroll
                            for (int j = 0; j < 8; ++j)
                                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + 4 * j),
                                             "f"(v[4 * j]), "f"(v[4 * j + 1]), "f"(v[4 * j + 2]), "f"(v[4 * j + 3])
                                             : "memory");
                        } else {
#pragma unroll
                            for (int j = 0; j < 8; ++j)
                                reinterpret_cast<float4*>(dst)[j] =
                                    make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                        }
                    }
                } else {
                    // ragged right edge or unaligned leading dimensions: scalar path
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const int gc = col0 + j;
                        if (gc >= p.N) continue;
                        float x = __uint_as_float(r[j]);
                        if (p.bias) x += p.bias[gc];
                        if (p.out_pre) p.out_pre[grow * p.ld_pre + gc] = __float2bfloat16(x);
                        if (p.act != UC2_ACT_NONE) {
                            const float a =
                                p.act == UC2_ACT_DGELU ? __bfloat162float(p.aux[grow * p.ld_aux + gc]) : 0.f;
                            x = apply_act(x, p.act, a);
                        }
                        if (p.residual)
                            x += p.res_f32 ? reinterpret_cast<const float*>(p.residual)[grow * p.ld_res + gc]
                                           : __bfloat162float(p.residual[grow * p.ld_res + gc]);
                        if (p.out_bf16) p.out_bf16[grow * p.ld_out + gc] = __float2bfloat16(x);
                        if (p.out_f32) {
                            if (p.accumulate) atomicAdd(p.out_f32 + grow * p.ld_f32 + gc, x);
                            else p.out_f32[grow * p.ld_f32 + gc] = x;
                        }
                    }
                }
            }
            if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
        }
    }

    // ------------------------------------------- teardown -------------------------------------------
    ptx::tc_fence_before();
    __syncthreads();
    if (warp_idx == 1) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc(tmem_base, C::TMEM_COLS);
    }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* f = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(f);
    });
    return fn;
}

// 2-D bf16 tensor [rows][cols] with row pitch ld elements; box = [box_rows][64 cols], SWIZZLE_128B.
int make_tmap(CUtensorMap* m, const void* base, long long rows, long long cols, long long ld, int box_rows) {
    EncodeTiledFn enc = get_encode_fn();
    UC2_REQUIRE(enc != nullptr, UC2_ERR_CUDA, "cuTensorMapEncodeTiled entry point not found");
    cuuint64_t gdim[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
    cuuint64_t gstr[1] = {static_cast<cuuint64_t>(ld) * 2};
    cuuint32_t box[2] = {64u, static_cast<cuuint32_t>(box_rows)};
    cuuint32_t estr[2] = {1u, 1u};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstr, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    UC2_REQUIRE(r == CUDA_SUCCESS, UC2_ERR_CUDA,
                "cuTensorMapEncodeTiled failed (%d): rows=%lld cols=%lld ld=%lld box_rows=%d base=%p", (int)r, rows,
                cols, ld, box_rows, base);
    return UC2_OK;
}

template <int BLOCK_N, bool A_MN, bool B_MN>
int launch(const uc2_gemm_args& a, const GemmParams& p, cudaStream_t stream) {
    using C = Cfg<BLOCK_N>;
    static_assert(C::STAGES >= 3, "pipeline too shallow");
    auto kern = gemm_bf16_kernel<BLOCK_N, A_MN, B_MN>;
    static std::once_flag once;
    static cudaError_t attr_err = cudaSuccess;
    std::call_once(once, [&] {
        attr_err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES);
    });
    UC2_REQUIRE(attr_err == cudaSuccess, UC2_ERR_CUDA, "cudaFuncSetAttribute(smem=%d): %s", C::SMEM_BYTES,
                cudaGetErrorString(attr_err));
    CUtensorMap ta, tb;
    int rc;
    if (!A_MN) rc = make_tmap(&ta, a.a, a.M, a.K, a.lda, BLOCK_M);
    else       rc = make_tmap(&ta, a.a, a.K, a.M, a.lda, BLOCK_K);
    if (rc) return rc;
    if (!B_MN) rc = make_tmap(&tb, a.b, a.N, a.K, a.ldb, BLOCK_N);
    else       rc = make_tmap(&tb, a.b, a.K, a.N, a.ldb, BLOCK_K);
    if (rc) return rc;
    const int total = p.num_m_blocks * p.num_n_blocks * p.split_k;
    const int grid = total < num_sms() ? total : num_sms();
    {
        ProfScope prof(stream, 0, 2.0 * a.M * a.N * a.K);
        kern<<<grid, GEMM_THREADS, C::SMEM_BYTES, stream>>>(ta, tb, p);
    }
    return check_last("gemm_bf16_kernel");
}

template <int BLOCK_N>
int dispatch_major(const uc2_gemm_args& a, const GemmParams& p, cudaStream_t s) {
    if (!a.a_mn && !a.b_mn) return launch<BLOCK_N, false, false>(a, p, s);
    if (!a.a_mn && a.b_mn) return launch<BLOCK_N, false, true>(a, p, s);
    if (a.a_mn && a.b_mn) return launch<BLOCK_N, true, true>(a, p, s);
    return launch<BLOCK_N, true, false>(a, p, s);
}

// Fewest waves, then least padding waste; ties -> the larger tile (fewer B re-reads).
int pick_block_n(int M, int N, int split_k) {
    const int sms = num_sms();
    const int mb = (M + BLOCK_M - 1) / BLOCK_M;
    double best = 1e30;
    int best_bn = 128;
    const int cands[3] = {256, 128, 64};
    for (int i = 0; i < 3; ++i) {
        const int bn = cands[i];
        const long long tiles = 1LL * mb * ((N + bn - 1) / bn) * split_k;
        const long long waves = (tiles + sms - 1) / sms;
        // per-tile cost ~ MMA time (prop. to bn) + fixed overhead; smaller N tiles pay relatively more
        const double cost = waves * (bn + 24.0);
        if (cost < best - 1e-9) { best = cost; best_bn = bn; }
    }
    return best_bn;
}

}  // namespace

}  // namespace uc2

extern "C" UC2_API int uc2_gemm_bf16(const uc2_gemm_args* args, void* stream) {
    using namespace uc2;
    UC2_REQUIRE(args != nullptr, UC2_ERR_ARG, "uc2_gemm_bf16: null args");
    const uc2_gemm_args& a = *args;
    if (int rc = require_sm100()) return rc;
    UC2_REQUIRE(a.M > 0 && a.N > 0 && a.K > 0, UC2_ERR_ARG, "uc2_gemm_bf16: bad shape %d %d %d", a.M, a.N, a.K);
    UC2_REQUIRE(a.a && a.b, UC2_ERR_ARG, "uc2_gemm_bf16: null operand");
    UC2_REQUIRE(aligned16(a.a) && aligned16(a.b), UC2_ERR_ARG, "uc2_gemm_bf16: operands must be 16-byte aligned");
    UC2_REQUIRE(a.lda % 8 == 0 && a.ldb % 8 == 0, UC2_ERR_ARG,
                "uc2_gemm_bf16: lda/ldb must be multiples of 8 elements (TMA 16-byte pitch), got %lld %lld", a.lda,
                a.ldb);
    UC2_REQUIRE(a.out_bf16 || a.out_f32, UC2_ERR_ARG, "uc2_gemm_bf16: no output");
    UC2_REQUIRE(a.act != UC2_ACT_DGELU || a.aux, UC2_ERR_ARG, "uc2_gemm_bf16: DGELU needs aux");
    int split_k = a.split_k < 0 ? 1 : a.split_k;
    if (split_k == 0) {
        // auto: only meaningful for the atomic-accumulate (wgrad) form; fill the SMs with K splits
        split_k = 1;
        const bool can_split = a.out_f32 && a.accumulate && !a.out_bf16 && !a.out_pre && !a.bias && !a.residual &&
                               a.act == UC2_ACT_NONE;
        if (can_split) {
            const int bn0 = a.block_n ? a.block_n : (a.N >= 256 ? 256 : (a.N >= 128 ? 128 : 64));
            const long long tiles = 1LL * ((a.M + BLOCK_M - 1) / BLOCK_M) * ((a.N + bn0 - 1) / bn0);
            const int kblocks = (a.K + BLOCK_K - 1) / BLOCK_K;
            const int sms = num_sms();
            double best = 0.0;
            for (int s = 1; s <= 32 && kblocks / s >= 4; ++s) {
                const long long t = tiles * s;
                const double eff = (double)t / (double)(((t + sms - 1) / sms) * sms);
                if (eff > best + 0.03) { best = eff; split_k = s; }
            }
        }
    }
    if (split_k > 1)
        UC2_REQUIRE(a.out_f32 && a.accumulate && !a.out_bf16 && !a.out_pre && !a.bias && !a.residual &&
                        a.act == UC2_ACT_NONE,
                    UC2_ERR_ARG, "uc2_gemm_bf16: split_k>1 requires accumulate into out_f32 only");
    GemmParams p;
    p.M = a.M; p.N = a.N; p.K = a.K;
    p.num_m_blocks = (a.M + BLOCK_M - 1) / BLOCK_M;
    p.num_k_blocks = (a.K + BLOCK_K - 1) / BLOCK_K;
    p.split_k = split_k > p.num_k_blocks ? p.num_k_blocks : split_k;
    p.k_blocks_per_split = (p.num_k_blocks + p.split_k - 1) / p.split_k;
    p.split_k = (p.num_k_blocks + p.k_blocks_per_split - 1) / p.k_blocks_per_split;   // no empty splits
    p.bias = a.bias;
    p.residual = static_cast<const bf16*>(a.residual); p.ld_res = a.ld_res; p.res_f32 = a.residual_f32;
    p.aux = static_cast<const bf16*>(a.aux); p.ld_aux = a.ld_aux;
    p.act = a.act;
    p.out_bf16 = static_cast<bf16*>(a.out_bf16); p.ld_out = a.ld_out;
    p.out_pre = static_cast<bf16*>(a.out_pre); p.ld_pre = a.ld_pre;
    p.out_f32 = a.out_f32; p.ld_f32 = a.ld_f32;
    p.accumulate = a.accumulate;
    auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
    p.vec_ok = (!a.residual || (a.ld_res % (a.residual_f32 ? 4 : 8) == 0 && al16(a.residual))) &&
               (!a.aux || (a.ld_aux % 8 == 0 && al16(a.aux))) &&
               (!a.out_bf16 || (a.ld_out % 8 == 0 && al16(a.out_bf16))) &&
               (!a.out_pre || (a.ld_pre % 8 == 0 && al16(a.out_pre))) &&
               (!a.out_f32 || (a.ld_f32 % 4 == 0 && al16(a.out_f32))) && (!a.bias || al16(a.bias));
    int bn = a.block_n;
    if (bn == 0) bn = pick_block_n(a.M, a.N, p.split_k);
    UC2_REQUIRE(bn == 64 || bn == 128 || bn == 256, UC2_ERR_ARG, "uc2_gemm_bf16: block_n must be 64/128/256");
    p.num_n_blocks = (a.N + bn - 1) / bn;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (bn == 256) return dispatch_major<256>(a, p, s);
    if (bn == 128) return dispatch_major<128>(a, p, s);
    return dispatch_major<64>(a, p, s);
}
