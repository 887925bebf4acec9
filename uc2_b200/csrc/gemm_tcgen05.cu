// Persistent, warp-specialised bf16 GEMM for sm_100a:
//   TMA (cp.async.bulk.tensor, SWIZZLE_128B) -> shared memory ring -> tcgen05.mma (fp32 accumulators
//   in TMEM, double buffered) -> tcgen05.ld epilogue with fused bias / GELU / dGELU / tanh / residual.
// One kernel template covers forward (A,B K-major), dgrad (B MN-major) and wgrad (A,B MN-major,
// split-K with fp32 atomic accumulation).  Replaces the nn.Linear call sites listed in
// include/uc2_b200.h (model/layer.py:76-78,112,140,153; model/model.py:359,1153-1169).
//
// CTAS = 2 runs the kernel as CTA pairs (thread-block clusters of two on one TPC, tcgen05 cta_group::2):
// a pair owns a 256 x BLOCK_N tile, each CTA stages its 128 rows of A and its half of the B rows, so a
// B byte is read from L2 and written to shared memory once per 256 output rows instead of once per 128
// (the 128-row form is L2- and shared-memory-bandwidth bound well below the tensor-pipe rate).
//
// Warp roles (320 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer (leader CTA),
// warps 2..9 = epilogue.  An epilogue warp owns TMEM lane quarter warp_idx % 4 (the hardware's rule) and
// every other 32-column chunk of the accumulator; a thread owns one output row of a chunk and moves it
// with 256-bit global accesses (one whole 32-byte sector per thread per instruction, nothing goes through
// shared memory, whose bandwidth belongs to the MMAs).  The residual / dGELU operand of a tile is requested
// BEFORE the accumulator is complete, so those loads fly while the MMAs of the tile still run, and the
// epilogue of tile i hides under the MMAs of tile i+1 (two accumulator buffers).
//
// Tile order: fixed round-robin over the persistent workers, or (uc2_gemm_sched_dynamic, on under data parallelism)
// drawn from a per-launch device counter and handed to the roles through a shared-memory ring -- see SCHED_R.
// The GELU / GELU' epilogues evaluate two columns per instruction on packed fp32 pairs (common.cuh: gelu_erf2).
#include <mutex>

#include "common.cuh"
#include "ptx.cuh"

namespace uc2 {

namespace {

constexpr int BLOCK_M = 128;   // rows per CTA; a CTA pair covers 2 x BLOCK_M
constexpr int BLOCK_K = 64;    // 64 bf16 = 128 B = one SWIZZLE_128B row
constexpr int UMMA_K = 16;
constexpr int EPI_COLS = 32;   // accumulator columns per tcgen05.ld (one fp32 row slice of 128 B per thread)
constexpr int SMEM_LIMIT = 232448;   // 227 KB opt-in maximum per CTA

// Epilogue warps per CTA: 8 (two per TMEM lane quarter, 168 registers each), or 16 for the two ALU-heavy
// epilogues (FFN1 + GELU with two outputs, FFN2 dgrad * gelu'): they need few registers (no fp32 operand
// prefetch) but twice the issue slots to keep up with K = 768 mainloops.
__host__ __device__ constexpr int epi_warps(int mode) { return (mode == 1 || mode == 2) ? 16 : 8; }
__host__ __device__ constexpr int gemm_threads(int mode) { return 64 + 32 * epi_warps(mode); }

// fp32-output epilogues (O-proj / FFN2 forward, modes 3 and 8) leave through TMA: a warp parks its 32 x 32 fp32
// chunk in a swizzled shared-memory box and one lane issues cp.async.bulk.tensor -- 32 rows x 128 B as one
// request instead of 64 row-scattered 32-byte sectors through the L1 tag stage, which is what bounded those two.
__host__ __device__ constexpr bool tma_out(int mode) { return mode == 3 || mode == 8; }
constexpr int OUT_BOX_BYTES = 32 * 32 * 4;

// Dynamic tile scheduler.  The persistent workers (CTAs, or CTA pairs) of a launch draw their tiles from one global
// counter instead of the fixed round-robin  t = worker, worker + n, ...  so that a worker which becomes resident late
// -- its SM was held by another kernel, typically a NCCL collective of the overlapped gradient exchange -- finds the
// work already done and leaves, instead of the whole GEMM waiting for a late starter to walk through its full share.
// One thread per worker (the TMA producer of the leader CTA) draws the index and hands it to the other roles of
// both CTAs through a small ring in shared memory: sched_full[slot] (count 1, in every CTA) says the slot holds tile
// number `it`, sched_empty[slot] (in the leader, one arrive per consumer of both CTAs) says everybody has read it.
constexpr int SCHED_R = 4;           // ring slots: the producer runs at most a tile or two ahead of the epilogue
constexpr int BAR_BYTES = 384;       // mbarriers + TMEM pointer + scheduler ring at the end of shared memory
constexpr int SCHED_SLOTS = 1024;    // (counter, finished workers) pairs handed out to launches in turn

template <int BLOCK_N, int CTAS, int MODE = 0>
struct Cfg {
    static constexpr int A_BYTES = BLOCK_M * BLOCK_K * 2;
    static constexpr int B_ROWS = BLOCK_N / CTAS;            // B rows (n) staged by one CTA
    static constexpr int B_BYTES = B_ROWS * BLOCK_K * 2;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int STAGING_BYTES = tma_out(MODE) ? epi_warps(MODE) * OUT_BOX_BYTES : 0;
    static constexpr int MAX_STAGES = (SMEM_LIMIT - 1024 - BAR_BYTES - STAGING_BYTES) / STAGE_BYTES;
    static constexpr int STAGES = MAX_STAGES > 8 ? 8 : MAX_STAGES;
    static constexpr int TMEM_COLS = 2 * BLOCK_N;   // two accumulator buffers; 128/256/512: power of two
    // layout: [stages][staging boxes (1024-byte aligned: STAGE_BYTES is a multiple of 1024)][barriers]
    static constexpr int SMEM_BYTES = 1024 + STAGES * STAGE_BYTES + STAGING_BYTES + BAR_BYTES;
};

struct GemmParams {
    int M, N, K;
    int num_m_blocks, num_n_blocks, split_k, k_blocks_per_split, num_k_blocks;   // m blocks of BLOCK_M * CTAS rows
    const float* bias;
    const bf16* residual; long long ld_res; int res_f32;
    const bf16* aux; long long ld_aux;
    int act;
    bf16* out_bf16; long long ld_out;
    bf16* out_pre; long long ld_pre;
    float* out_f32; long long ld_f32;
    int accumulate;
    int vec_ok;   // leading dimensions / pointers allow 32-byte row-slice access -> vector epilogue
    DropCfg drop; // dropout on (acc + bias) before the residual add (BertSelfOutput / BertOutput, layer.py:113,154)
    // Tail split: tiles [0, tail_start) are full BLOCK_N-wide tiles; each of the remaining tiles_mn - tail_start
    // tiles (the partial last round of the persistent workers) is cut into tail_split column slices so that round
    // costs ~1/tail_split of a full one (19200 tokens = 75 row blocks on 74 CTA pairs would otherwise pay a whole
    // extra round for 1-4 % of the work).  tail_split == 1: off.
    int tail_start, tail_split;
    // MODE 9: cross-entropy statistics per (row, 32-column chunk), see uc2_gemm_args
    float2* ce_stats; long long ce_ld; const long long* ce_labels; float* ce_tgt;
    // dynamic tile scheduler: sched[0] = next tile, sched[1] = workers that have drawn their last (invalid) tile; both
    // zero at launch, the last worker to finish zeroes them again for the launch that gets this pair next
    int* sched;
};

struct TileCoord { int m_blk, n0, width, split; };

template <int BLOCK_N>
__device__ __forceinline__ TileCoord decode_tile(const GemmParams& p, int t, int tiles_mn) {
    TileCoord c;
    int base = t;
    c.width = BLOCK_N;
    int sub = 0;
    if (p.tail_split > 1 && t >= p.tail_start) {
        const int r = t - p.tail_start;
        base = p.tail_start + r / p.tail_split;
        sub = r % p.tail_split;
        c.width = BLOCK_N / p.tail_split;
    }
    c.split = base / tiles_mn;
    const int mn = base % tiles_mn;
    c.m_blk = mn / p.num_n_blocks;
    c.n0 = (mn % p.num_n_blocks) * BLOCK_N + sub * c.width;
    return c;
}

__device__ __forceinline__ float apply_act(float v, int act, float aux) {
    if (act == UC2_ACT_GELU) return gelu_erf(v);
    if (act == UC2_ACT_DGELU) return v * gelu_erf_grad(aux);
    if (act == UC2_ACT_TANH) return fast_tanh(v);
    return v;
}

__device__ __forceinline__ void store_bf16_row16(bf16* dst, const float* v) {
    ptx::stg256(dst, pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]),
                pack_bf16(v[8], v[9]), pack_bf16(v[10], v[11]), pack_bf16(v[12], v[13]), pack_bf16(v[14], v[15]));
}

// Epilogue specialisation.  MODE 0 reads every switch from GemmParams at run time; MODE 1..7 are the seven
// epilogues of the BertLayer stack with the switches fixed at compile time (no branches, a fraction of the code:
// the epilogue warps were stalling on instruction fetch and uniform branches in the generic form).
//   1 FFN1 forward   : + bias, pre-activation copy, erf-GELU            -> bf16
//   2 FFN2 dgrad     : * gelu'(aux)                                      -> bf16
//   3 O-proj / FFN2  : + bias + fp32 residual                            -> fp32
//   4 QKV forward    : + bias                                            -> bf16
//   5 FFN1/QKV dgrad : + bf16 residual                                   -> bf16
//   6 wgrad          : fp32 atomic accumulate
//   7 O-proj dgrad   : plain                                             -> bf16
//   8 O-proj / FFN2  : dropout(+ bias) + fp32 residual (training)        -> fp32
//   9 MLM decoder    : + bias -> bf16 logits, and (max, sum-exp) per row and 32-column chunk + the target logit
template <int MODE>
struct Epi {
    static __device__ __forceinline__ bool bias(const GemmParams& p) {
        return MODE == 0 ? p.bias != nullptr : (MODE == 1 || MODE == 3 || MODE == 4 || MODE == 8 || MODE == 9);
    }
    static __device__ __forceinline__ bool out_pre(const GemmParams& p) { return MODE == 0 ? p.out_pre != nullptr : MODE == 1; }
    static __device__ __forceinline__ int act(const GemmParams& p) {
        return MODE == 0 ? p.act : (MODE == 1 ? UC2_ACT_GELU : (MODE == 2 ? UC2_ACT_DGELU : UC2_ACT_NONE));
    }
    static __device__ __forceinline__ bool residual(const GemmParams& p) {
        return MODE == 0 ? p.residual != nullptr : (MODE == 3 || MODE == 5 || MODE == 8);
    }
    static __device__ __forceinline__ bool out_bf16(const GemmParams& p) {
        return MODE == 0 ? p.out_bf16 != nullptr
                         : (MODE == 1 || MODE == 2 || MODE == 4 || MODE == 5 || MODE == 7 || MODE == 9);
    }
    static __device__ __forceinline__ bool out_f32(const GemmParams& p) {
        return MODE == 0 ? p.out_f32 != nullptr : (MODE == 3 || MODE == 6 || MODE == 8);
    }
    static __device__ __forceinline__ bool accumulate(const GemmParams& p) { return MODE == 0 ? p.accumulate != 0 : MODE == 6; }
    static __device__ __forceinline__ bool dropout(const GemmParams& p) { return MODE == 0 ? p.drop.thresh != 0 : MODE == 8; }
    static constexpr int ex_kind = (MODE == 2 || MODE == 5) ? 1 : ((MODE == 3 || MODE == 8) ? 2 : 0);   // for MODE != 0
};

__device__ __noinline__ void epilogue_scalar_row(const GemmParams& p, const float* acc, long long grow, int col0) {
#pragma unroll 1
    for (int j = 0; j < 16; ++j) {
        const int gc = col0 + j;
        if (gc >= p.N) break;
        float x = acc[j];
        if (p.bias) x += p.bias[gc];
        if (p.out_pre) p.out_pre[grow * p.ld_pre + gc] = __float2bfloat16(x);
        if (p.act != UC2_ACT_NONE) {
            const float a = p.act == UC2_ACT_DGELU ? __bfloat162float(p.aux[grow * p.ld_aux + gc]) : 0.f;
            x = apply_act(x, p.act, a);
        }
        if (p.drop.thresh != 0)
            x = drop_keep(p.drop.key, static_cast<uint32_t>(grow) * static_cast<uint32_t>(p.N) + gc, p.drop.thresh)
                    ? x * p.drop.scale : 0.f;
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

// One accumulator tile (this warp's 32 rows x every other 32-column chunk) through the fused epilogue.
// EX: extra operand read by the epilogue (0 none, 1 bf16 residual or dGELU aux, 2 fp32 residual).  Chunks go in
// groups of two: the group's operand loads are issued together (for the first group: before the accumulator is
// complete, so they fly while the MMAs of the tile still run), then each chunk is processed 16 columns at a
// time, which keeps the live registers low enough that nothing spills (local memory has no L1 behind it here:
// the shared-memory carve-out is the whole 227 KB).
template <int BLOCK_N, int EX, int CTAS, int MODE>
__device__ __forceinline__ void epilogue_tile(const GemmParams& p, uint32_t tacc, int lane, int half, long long grow,
                                              int ncol0, int ncols, uint32_t tfull, uint32_t tfull_phase,
                                              uint32_t tempty, const CUtensorMap* tmap_out = nullptr,
                                              uint32_t out_box = 0) {
    using F = Epi<MODE>;
    constexpr int CG = epi_warps(MODE) / 4;          // warps sharing a lane quarter; `half` = this warp's index there
    static_assert(BLOCK_N / EPI_COLS >= CG, "tile too narrow for this many epilogue warps");
    constexpr int MY = BLOCK_N / EPI_COLS / CG;      // chunks per warp (chunk c = CG * i + half)
    constexpr int G = MY < 2 ? MY : 2;               // chunks per group
    constexpr int W = EX == 2 ? 32 : 16;             // 32-bit words of the extra operand per chunk row
    const bool row_ok = grow < p.M;
    // MODE 9: running (max, sum exp) of this row over the two halves of a chunk, and the row's label
    float ce_m = 0.f, ce_s = 0.f;
    const int ce_t = (MODE == 9 && row_ok) ? static_cast<int>(p.ce_labels[grow]) : -1;
    uint32_t pf[EX == 0 ? 1 : G * W];
    // MODE 1: the 16 bias values of the NEXT 16-column slice are requested while the current slice goes through the
    // GELU (the first slice of a tile: before the wait for the accumulator) -- the shared-memory carve-out leaves
    // almost no L1, so a bias load issued at its point of use is a trip to L2 with the warp stalled on it
    float4 bq[MODE == 1 ? 4 : 1];
    auto bias_ok = [&](int c) { return p.vec_ok && c * EPI_COLS < ncols && ncol0 + (c + 1) * EPI_COLS <= p.N; };
    auto bias_fetch = [&](int c, int h) {
#pragma unroll
        for (int j = 0; j < (MODE == 1 ? 4 : 1); ++j)
            bq[j] = __ldg(reinterpret_cast<const float4*>(p.bias + ncol0 + c * EPI_COLS + h * 16) + j);
    };
    const bf16* exb = F::residual(p) ? p.residual : p.aux;
    const long long ld_ex = F::residual(p) ? p.ld_res : p.ld_aux;
#pragma unroll 1
    for (int g0 = 0; g0 < MY; g0 += G) {
        if (EX != 0) {
#pragma unroll
            for (int ii = 0; ii < G; ++ii) {
                const int col0 = ncol0 + (CG * (g0 + ii) + half) * EPI_COLS;
                if (p.vec_ok && col0 + EPI_COLS <= p.N && row_ok && (CG * (g0 + ii) + half) * EPI_COLS < ncols) {
                    const uint8_t* src = EX == 2
                        ? reinterpret_cast<const uint8_t*>(reinterpret_cast<const float*>(exb) + grow * ld_ex + col0)
                        : reinterpret_cast<const uint8_t*>(exb + grow * ld_ex + col0);
#pragma unroll
                    for (int j = 0; j < W / 8; ++j) ptx::ldg256(src + 32 * j, pf + ii * W + 8 * j);
                }
            }
        }
        if (g0 == 0) {
            if (MODE == 1 && row_ok && bias_ok(half)) bias_fetch(half, 0);
            ptx::mbar_wait(tfull, tfull_phase);
            ptx::tc_fence_after();
        }
#pragma unroll
        for (int ii = 0; ii < G; ++ii) {
            const int c = CG * (g0 + ii) + half;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int col0 = ncol0 + c * EPI_COLS + h * 16;
                const bool in_tile = c * EPI_COLS < ncols;          // narrow tail tiles use the first columns only
                uint32_t r[16];
                if (in_tile) {
                    ptx::tmem_ld_32x16(tacc + c * EPI_COLS + h * 16, r);
                    ptx::tmem_wait_ld();
                }
                if (g0 + ii == MY - 1 && h == 1) {
                    // all TMEM reads of this accumulator are done: hand it back to the MMA warp (relaxed arrive:
                    // nothing in generic memory is published here, and a release at cluster scope would make the
                    // warp wait for all of its outstanding global stores)
                    ptx::tc_fence_before();
                    __syncwarp();
                    if (lane == 0) {
                        if (CTAS == 2) ptx::mbar_arrive_remote_relaxed(tempty);
                        else ptx::mbar_arrive(tempty);
                    }
                }
                if (!in_tile || col0 >= p.N) continue;
                // the TMA-store path needs the whole warp (rows beyond M are clipped by the tensor map)
                if (!row_ok && !tma_out(MODE)) continue;
                const uint32_t* ex = pf + ii * W + h * (W / 2);
                if ((MODE == 1 || MODE == 2) && p.vec_ok && ncol0 + (c + 1) * EPI_COLS <= p.N) {
                    // the two GELU epilogues on packed fp32 pairs (common.cuh: FFMA2 / FMUL2 / FADD2), same values
                    f32x2 w[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) w[j] = pk2u(r[2 * j], r[2 * j + 1]);
                    if (MODE == 1) {
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            w[2 * j] = add2(w[2 * j], pk2(bq[j].x, bq[j].y));
                            w[2 * j + 1] = add2(w[2 * j + 1], pk2(bq[j].z, bq[j].w));
                        }
                        if (h == 0) {
                            bias_fetch(c, 1);
                        } else {
                            const int cn = CG * (g0 + ii + 1) + half;
                            if (g0 + ii + 1 < MY && bias_ok(cn)) bias_fetch(cn, 0);
                        }
                        ptx::stg256(p.out_pre + grow * p.ld_pre + col0, pack_bf16x2(w[0]), pack_bf16x2(w[1]),
                                    pack_bf16x2(w[2]), pack_bf16x2(w[3]), pack_bf16x2(w[4]), pack_bf16x2(w[5]),
                                    pack_bf16x2(w[6]), pack_bf16x2(w[7]));
#pragma unroll
                        for (int j = 0; j < 8; ++j) w[j] = gelu_erf2(w[j]);
                    } else {
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const float2 u = unpack_bf16(ex[j]);
                            w[j] = mul2(w[j], gelu_erf_grad2(pk2(u.x, u.y)));
                        }
                    }
                    ptx::stg256(p.out_bf16 + grow * p.ld_out + col0, pack_bf16x2(w[0]), pack_bf16x2(w[1]),
                                pack_bf16x2(w[2]), pack_bf16x2(w[3]), pack_bf16x2(w[4]), pack_bf16x2(w[5]),
                                pack_bf16x2(w[6]), pack_bf16x2(w[7]));
                } else if (p.vec_ok && ncol0 + (c + 1) * EPI_COLS <= p.N) {
                    float v[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]);
                    if (F::bias(p)) {
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + col0) + j);
                            v[4 * j] += b.x; v[4 * j + 1] += b.y; v[4 * j + 2] += b.z; v[4 * j + 3] += b.w;
                        }
                    }
                    if (MODE == 9) {
                        constexpr float L2E = 1.4426950408889634f;
                        float cm = v[0];
#pragma unroll
                        for (int j = 1; j < 16; ++j) cm = fmaxf(cm, v[j]);
                        const float nm = h == 0 ? cm : fmaxf(ce_m, cm), nml = nm * L2E;
                        float ss = 0.f;
#pragma unroll
                        for (int j = 0; j < 16; ++j) ss += fast_ex2(fmaf(v[j], L2E, -nml));
                        ce_s = h == 0 ? ss : fmaf(ce_s, fast_ex2(fmaf(ce_m, L2E, -nml)), ss);
                        ce_m = nm;
                        if (static_cast<unsigned>(ce_t - col0) < 16u) {          // the label's own logit, in fp32
                            float z = v[0];
#pragma unroll
                            for (int j = 1; j < 16; ++j) z = (ce_t - col0 == j) ? v[j] : z;
                            p.ce_tgt[grow] = z;
                        }
                        if (h == 1) p.ce_stats[static_cast<long long>(col0 >> 5) * p.ce_ld + grow] = make_float2(ce_m, ce_s);
                    }
                    if (F::out_pre(p)) store_bf16_row16(p.out_pre + grow * p.ld_pre + col0, v);
                    if (F::act(p) == UC2_ACT_GELU) {
#pragma unroll
                        for (int j = 0; j < 16; ++j) v[j] = gelu_erf(v[j]);
                    } else if (F::act(p) == UC2_ACT_TANH) {
#pragma unroll
                        for (int j = 0; j < 16; ++j) v[j] = fast_tanh(v[j]);
                    } else if (F::act(p) == UC2_ACT_DGELU) {
                        if (EX == 1 && !F::residual(p)) {
#pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                const float2 u = unpack_bf16(ex[j]);
                                v[2 * j] *= gelu_erf_grad(u.x);
                                v[2 * j + 1] *= gelu_erf_grad(u.y);
                            }
                        } else {
                            // residual and aux together (not used by the encoder): aux fetched late
                            const uint32_t* a = reinterpret_cast<const uint32_t*>(p.aux + grow * p.ld_aux + col0);
#pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                const float2 u = unpack_bf16(__ldg(a + j));
                                v[2 * j] *= gelu_erf_grad(u.x);
                                v[2 * j + 1] *= gelu_erf_grad(u.y);
                            }
                        }
                    }
                    if (F::dropout(p)) {
                        const uint32_t i0 = static_cast<uint32_t>(grow) * static_cast<uint32_t>(p.N) + col0;
#pragma unroll
                        for (int j = 0; j < 16; j += 2) {
                            bool k0, k1;
                            drop_keep2(p.drop.key, i0 + j, p.drop.thresh, k0, k1);
                            v[j] = k0 ? v[j] * p.drop.scale : 0.f;
                            v[j + 1] = k1 ? v[j + 1] * p.drop.scale : 0.f;
                        }
                    }
                    if (EX == 2) {
#pragma unroll
                        for (int j = 0; j < 16; ++j) v[j] += __uint_as_float(ex[j]);
                    } else if (EX == 1 && F::residual(p)) {
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const float2 u = unpack_bf16(ex[j]);
                            v[2 * j] += u.x;
                            v[2 * j + 1] += u.y;
                        }
                    }
                    if (F::out_bf16(p)) store_bf16_row16(p.out_bf16 + grow * p.ld_out + col0, v);
                    if (tma_out(MODE)) {
                        // park this row's 16 values in the swizzled box (16-byte unit u of row r sits at u ^ (r & 7))
                        if (h == 0) {
                            if (lane == 0) ptx::tma_store_wait_read<0>();    // the previous chunk's store has read the box
                            __syncwarp();
                        }
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            ptx::st_shared_v4(out_box + lane * 128 + (((h * 4 + j) ^ (lane & 7)) << 4),
                                              __float_as_uint(v[4 * j]), __float_as_uint(v[4 * j + 1]),
                                              __float_as_uint(v[4 * j + 2]), __float_as_uint(v[4 * j + 3]));
                        if (h == 1) {
                            ptx::fence_proxy_async();
                            __syncwarp();
                            if (lane == 0) {
                                ptx::tma_store_2d(tmap_out, out_box, ncol0 + c * EPI_COLS,
                                                  static_cast<int>(grow - lane));
                                ptx::tma_store_commit();
                            }
                        }
                    } else if (F::out_f32(p)) {
                        float* dst = p.out_f32 + grow * p.ld_f32 + col0;
                        if (F::accumulate(p)) {
#pragma unroll
                            for (int j = 0; j < 4; ++j)
                                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + 4 * j),
                                             "f"(v[4 * j]), "f"(v[4 * j + 1]), "f"(v[4 * j + 2]), "f"(v[4 * j + 3])
                                             : "memory");
                        } else {
#pragma unroll
                            for (int j = 0; j < 2; ++j)
                                ptx::stg256(dst + 8 * j, __float_as_uint(v[8 * j]), __float_as_uint(v[8 * j + 1]),
                                            __float_as_uint(v[8 * j + 2]), __float_as_uint(v[8 * j + 3]),
                                            __float_as_uint(v[8 * j + 4]), __float_as_uint(v[8 * j + 5]),
                                            __float_as_uint(v[8 * j + 6]), __float_as_uint(v[8 * j + 7]));
                        }
                    }
                } else {
                    // ragged right edge or unaligned leading dimensions: scalar path (out of line, rolled up: it
                    // runs for at most one chunk per row of tiles on the shapes of this model)
                    float loc[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) loc[j] = __uint_as_float(r[j]);
                    if (MODE == 9) {
                        // the chunk that holds column N - 1: same outputs, element by element
                        if (row_ok) {
#pragma unroll 1
                            for (int j = 0; j < 16 && col0 + j < p.N; ++j) {
                                const float z = loc[j] + p.bias[col0 + j];
                                p.out_bf16[grow * p.ld_out + col0 + j] = __float2bfloat16(z);
                                if (h == 0 && j == 0) {
                                    ce_m = z;
                                    ce_s = 1.f;
                                } else {
                                    const float nm = fmaxf(ce_m, z);
                                    ce_s = ce_s * __expf(ce_m - nm) + __expf(z - nm);
                                    ce_m = nm;
                                }
                                if (col0 + j == ce_t) p.ce_tgt[grow] = z;
                            }
                            if (h == 1 || col0 + 16 >= p.N)
                                p.ce_stats[static_cast<long long>(col0 >> 5) * p.ce_ld + grow] = make_float2(ce_m, ce_s);
                        }
                    } else if (row_ok) {
                        epilogue_scalar_row(p, loc, grow, col0);
                    }
                }
            }
        }
    }
}

template <int BLOCK_N, bool A_MN, bool B_MN, int CTAS, int MODE>
__global__ void __launch_bounds__(gemm_threads(MODE), 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                 const __grid_constant__ CUtensorMap tmap_b_tail, const __grid_constant__ CUtensorMap tmap_out,
                 const GemmParams p) {
    using C = Cfg<BLOCK_N, CTAS, MODE>;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem_gen = smem_raw + (smem_base - ptx::smem_u32(smem_raw));
    const uint32_t bar_base = smem_base + C::STAGES * C::STAGE_BYTES + C::STAGING_BYTES;
    // barrier layout (8 B each): full[STAGES], empty[STAGES], tmem_full[2], tmem_empty[2], tmem ptr,
    // sched_full[SCHED_R], sched_empty[SCHED_R], then the ring of tile numbers (4 B each)
    static_assert(8 * (2 * C::STAGES + 5 + 2 * SCHED_R) + 4 * SCHED_R <= BAR_BYTES, "barrier region too small");
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (C::STAGES + s); };
    auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * C::STAGES + a); };
    auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * C::STAGES + 2 + a); };
    const uint32_t tmem_ptr_addr = bar_base + 8u * (2 * C::STAGES + 4);
    auto sfull_bar = [&](int s) { return bar_base + 8u * (2 * C::STAGES + 5 + s); };
    auto sempty_bar = [&](int s) { return bar_base + 8u * (2 * C::STAGES + 5 + SCHED_R + s); };
    auto ring_slot = [&](int s) { return bar_base + 8u * (2 * C::STAGES + 5 + 2 * SCHED_R) + 4u * s; };
    volatile uint32_t* tmem_ptr_gen =
        reinterpret_cast<volatile uint32_t*>(smem_gen + C::STAGES * C::STAGE_BYTES + C::STAGING_BYTES +
                                             8 * (2 * C::STAGES + 4));

    const int warp_idx = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t cta_rank = CTAS == 2 ? ptx::cluster_ctarank() : 0u;
    const int unit = CTAS == 2 ? (blockIdx.x >> 1) : blockIdx.x;          // persistent worker (CTA or CTA pair)
    const int num_units = CTAS == 2 ? (gridDim.x >> 1) : gridDim.x;
    // p.sched == nullptr: fixed round-robin tiles (t = unit, unit + num_units, ...); otherwise drawn from a counter.
    // The first number is requested before anything else so that the atomic's round trip runs under the prologue.
    const bool dyn = p.sched != nullptr;
    int first_tile = unit;
    if (dyn && threadIdx.x == 0 && cta_rank == 0) first_tile = atomicAdd(p.sched, 1);

    if (threadIdx.x == 0) {
        ptx::prefetch_tmap(&tmap_a);
        ptx::prefetch_tmap(&tmap_b);
        if (tma_out(MODE)) ptx::prefetch_tmap(&tmap_out);
        for (int s = 0; s < C::STAGES; ++s) {
            ptx::mbar_init(full_bar(s), 1);
            ptx::mbar_init(empty_bar(s), 1);
        }
        for (int a = 0; a < 2; ++a) {
            ptx::mbar_init(tfull_bar(a), 1);
            ptx::mbar_init(tempty_bar(a), epi_warps(MODE) * CTAS);   // one arrive per epilogue warp of every CTA
        }
        for (int s = 0; s < SCHED_R; ++s) {
            ptx::mbar_init(sfull_bar(s), 1);
            // readers of a tile number: the MMA thread and the leader's epilogue warps, plus (pairs) the peer's
            // producer and epilogue warps; the leader's producer is the writer
            ptx::mbar_init(sempty_bar(s), (1 + epi_warps(MODE)) * CTAS);
        }
        ptx::fence_barrier_init();
        ptx::fence_proxy_async();
    }
    if (warp_idx == 1) {
        if (CTAS == 2) {
            ptx::tmem_alloc_pair(tmem_ptr_addr, C::TMEM_COLS);
            ptx::tmem_relinquish_pair();
        } else {
            ptx::tmem_alloc(tmem_ptr_addr, C::TMEM_COLS);
            ptx::tmem_relinquish();
        }
    }
    ptx::tc_fence_before();
    if (CTAS == 2) ptx::cluster_sync();      // the peer's barriers must exist before anything signals them
    else __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_gen;

    const int tiles_mn = p.num_m_blocks * p.num_n_blocks;
    const int total_tiles = p.tail_split > 1 ? p.tail_start + (tiles_mn - p.tail_start) * p.tail_split
                                             : tiles_mn * p.split_k;

    // ---- tile scheduler (see SCHED_R) ----
    // writer side: the leader's producer thread.  The counter pair belongs to this launch alone, so drawing from it
    // does not have to wait for the previous kernel of the stream.
    auto sched_account = [&](int t) {
        if (t >= total_tiles && atomicAdd(p.sched + 1, 1) == num_units - 1) {
            // every worker has drawn its last number: leave the pair zeroed for the launch that gets it next
            atomicExch(p.sched, 0);
            atomicExch(p.sched + 1, 0);
        }
    };
    auto sched_publish = [&](int it, int t) {
        const int slot = it % SCHED_R;
        const uint32_t parity = (((uint32_t)it / SCHED_R) & 1u) ^ 1u;
        if (CTAS == 2) ptx::mbar_wait_cluster(sempty_bar(slot), parity);
        else ptx::mbar_wait(sempty_bar(slot), parity);
        ptx::st_shared_b32(ring_slot(slot), (uint32_t)t);
        ptx::mbar_arrive(sfull_bar(slot));
        if (CTAS == 2) {
            ptx::st_shared_cluster_b32(ptx::map_to_cta(ring_slot(slot), 1), (uint32_t)t);
            ptx::mbar_arrive_remote(ptx::map_to_cta(sfull_bar(slot), 1));       // release at cluster scope
        }
    };
    // reader side: waits for tile number `it` of this worker and returns it (>= total_tiles: no more work)
    auto sched_read = [&](int it) -> int {
        if (!dyn) return unit + it * num_units;
        const int slot = it % SCHED_R;
        const uint32_t parity = ((uint32_t)it / SCHED_R) & 1u;
        if (CTAS == 2) ptx::mbar_wait_cluster(sfull_bar(slot), parity);
        else ptx::mbar_wait(sfull_bar(slot), parity);
        return (int)ptx::ld_shared_b32(ring_slot(slot));
    };
    // ... and gives the slot back.  RELAXED: for the epilogue warps, whose release would wait for their global stores
    // in flight; the value has been consumed (shuffled) by then.  Otherwise a release arrive.
    auto sched_done = [&](int it, bool relaxed) {
        if (!dyn) return;
        const int slot = it % SCHED_R;
        if (CTAS == 2) {
            const uint32_t b = ptx::map_to_cta(sempty_bar(slot), 0);
            if (relaxed) ptx::mbar_arrive_remote_relaxed(b);
            else ptx::mbar_arrive_remote(b);
        } else {
            ptx::mbar_arrive(sempty_bar(slot));
        }
    };
    if (dyn && threadIdx.x == 0 && cta_rank == 0) {
        sched_account(first_tile);
        sched_publish(0, first_tile);
    }
    // everything above (barriers, TMEM, descriptor prefetch, the first tile number) may have run under the
    // previous kernel's tail
    griddep_sync();

    if (warp_idx == 0) {
        // ===================================== TMA producer =====================================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            const bool sched_writer = dyn && cta_rank == 0;
            int t = first_tile;
            for (int it = 0;; ++it) {
                if (!sched_writer) {
                    t = sched_read(it);
                    sched_done(it, false);
                }
                if (t >= total_tiles) break;
                int t_next = 0;
                const TileCoord tc = decode_tile<BLOCK_N>(p, t, tiles_mn);
                const int kb0 = tc.split * p.k_blocks_per_split;
                const int kb1 = min(kb0 + p.k_blocks_per_split, p.num_k_blocks);
                // the next tile number is requested while the last k blocks of this tile are still being issued and
                // looked at only after the last of them: the atomic's round trip never stalls the loads
                const int kb_draw = max(kb0, kb1 - 8);
                const int m0 = (tc.m_blk * CTAS + (int)cta_rank) * BLOCK_M;
                const int b_rows = tc.width / CTAS;                       // B rows (n) this CTA stages for the tile
                const int n0 = tc.n0 + (int)cta_rank * b_rows;
                const bool narrow = tc.width != BLOCK_N;
                const uint32_t stage_tx = (C::A_BYTES + b_rows * BLOCK_K * 2) * CTAS;
                for (int kb = kb0; kb < kb1; ++kb) {
                    if (sched_writer && kb == kb_draw) t_next = atomicAdd(p.sched, 1);
                    ptx::mbar_wait(empty_bar(stage), phase ^ 1u);
                    const uint32_t sa = smem_base + stage * C::STAGE_BYTES;
                    const uint32_t sb = sa + C::A_BYTES;
                    // the pair's loads all report to the leader's barrier, which expects both CTAs' bytes
                    const uint32_t fb = CTAS == 2 ? ptx::map_to_cta(full_bar(stage), 0) : full_bar(stage);
                    if (cta_rank == 0) ptx::mbar_arrive_expect_tx(full_bar(stage), stage_tx);
                    auto load = [&](uint32_t dst, const CUtensorMap* m, int c0, int c1) {
                        if (CTAS == 2) ptx::tma_load_2d_pair(dst, m, fb, c0, c1);
                        else ptx::tma_load_2d(dst, m, fb, c0, c1);
                    };
                    if (!A_MN) {
                        load(sa, &tmap_a, kb * BLOCK_K, m0);
                    } else {
#pragma unroll
                        for (int j = 0; j < BLOCK_M / 64; ++j) load(sa + j * 8192, &tmap_a, m0 + j * 64, kb * BLOCK_K);
                    }
                    if (!B_MN) {
                        load(sb, narrow ? &tmap_b_tail : &tmap_b, kb * BLOCK_K, n0);
                    } else {
#pragma unroll
                        for (int j = 0; j < C::B_ROWS / 64; ++j)
                            if (j * 64 < b_rows) load(sb + j * 8192, &tmap_b, n0 + j * 64, kb * BLOCK_K);
                    }
                    if (++stage == C::STAGES) { stage = 0; phase ^= 1u; }
                }
                if (sched_writer) {
                    sched_account(t_next);
                    sched_publish(it + 1, t_next);
                    t = t_next;
                }
            }
        }
    } else if (warp_idx == 1) {
        // ===================================== MMA issuer (leader CTA) ==========================
        if (lane == 0 && cta_rank == 0) {
            constexpr uint32_t idesc = ptx::idesc_bf16_f32(BLOCK_M * CTAS, BLOCK_N, A_MN, B_MN);
            // K-major: 8-row groups 1024 B apart (SBO); MN-major: 64-element blocks 8192 B apart (LBO),
            // 8-k-row groups 1024 B apart (SBO)
            constexpr uint32_t A_LBO = A_MN ? 8192u : 16u, B_LBO = B_MN ? 8192u : 16u;
            constexpr uint32_t A_KSTEP = A_MN ? 2048u : 32u, B_KSTEP = B_MN ? 2048u : 32u;
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            int t = sched_read(0);
            sched_done(0, true);
            for (int it = 0; t < total_tiles; ++it) {
                int t_next = 0;
                const TileCoord tc = decode_tile<BLOCK_N>(p, t, tiles_mn);
                const int kb0 = tc.split * p.k_blocks_per_split;
                const int kb1 = min(kb0 + p.k_blocks_per_split, p.num_k_blocks);
                const uint32_t idesc_t = tc.width == BLOCK_N ? idesc
                                                             : ptx::idesc_bf16_f32(BLOCK_M * CTAS, tc.width, A_MN, B_MN);
                if (CTAS == 2) ptx::mbar_wait_cluster(tempty_bar(acc), acc_phase ^ 1u);
                else ptx::mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
                ptx::tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
                for (int kb = kb0; kb < kb1; ++kb) {
                    if (kb == kb1 - 1) {
                        // the next tile number (published long ago: the producer is stages ahead) is fetched while the
                        // MMAs of the previous k block are still queued, not between two tiles
                        t_next = sched_read(it + 1);
                        sched_done(it + 1, true);
                    }
                    ptx::mbar_wait(full_bar(stage), phase);
                    ptx::tc_fence_after();
                    const uint32_t sa = smem_base + stage * C::STAGE_BYTES;
                    const uint32_t sb = sa + C::A_BYTES;
#pragma unroll
                    for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
                        const uint64_t da = ptx::smem_desc_sw128(sa + k * A_KSTEP, A_LBO, 1024u);
                        const uint64_t db = ptx::smem_desc_sw128(sb + k * B_KSTEP, B_LBO, 1024u);
                        const uint32_t accum = (kb > kb0 || k > 0) ? 1u : 0u;
                        if (CTAS == 2) ptx::umma_bf16_pair(d_tmem, da, db, idesc_t, accum);
                        else ptx::umma_bf16(d_tmem, da, db, idesc_t, accum);
                    }
                    // frees the smem slot (in both CTAs of a pair) once these MMAs retire
                    if (CTAS == 2) ptx::umma_commit_pair(empty_bar(stage));
                    else ptx::umma_commit(empty_bar(stage));
                    if (++stage == C::STAGES) { stage = 0; phase ^= 1u; }
                }
                // accumulator complete -> epilogue warps (of both CTAs)
                if (CTAS == 2) ptx::umma_commit_pair(tfull_bar(acc));
                else ptx::umma_commit(tfull_bar(acc));
                if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
                t = t_next;
            }
        }
    } else {
        // ===================================== epilogue =========================================
        const int q = warp_idx & 3;                       // TMEM lane quarter owned by this warp
        const int half = (warp_idx - 2) >> 2;             // which of the interleaved 32-column chunks
        const int ex_kind = MODE != 0 ? Epi<MODE>::ex_kind
                                      : (p.residual ? (p.res_f32 ? 2 : 1) : (p.act == UC2_ACT_DGELU ? 1 : 0));
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int it = 0;; ++it) {
            const int t = __shfl_sync(0xffffffffu, sched_read(it), 0);
            if (lane == 0) sched_done(it, true);
            if (t >= total_tiles) break;
            const TileCoord tc = decode_tile<BLOCK_N>(p, t, tiles_mn);
            const uint32_t tacc = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BLOCK_N;
            const long long grow = (long long)(tc.m_blk * CTAS + (int)cta_rank) * BLOCK_M + q * 32 + lane;
            const int ncol0 = tc.n0;
            const int ncols = tc.width;
            const uint32_t te = CTAS == 2 ? ptx::map_to_cta(tempty_bar(acc), 0) : tempty_bar(acc);
            if (MODE != 0)
                epilogue_tile<BLOCK_N, Epi<MODE>::ex_kind, CTAS, MODE>(
                    p, tacc, lane, half, grow, ncol0, ncols, tfull_bar(acc), acc_phase, te, &tmap_out,
                    smem_base + C::STAGES * C::STAGE_BYTES + (warp_idx - 2) * OUT_BOX_BYTES);
            else if (ex_kind == 0)
                epilogue_tile<BLOCK_N, 0, CTAS, 0>(p, tacc, lane, half, grow, ncol0, ncols, tfull_bar(acc), acc_phase, te);
            else if (ex_kind == 1)
                epilogue_tile<BLOCK_N, 1, CTAS, 0>(p, tacc, lane, half, grow, ncol0, ncols, tfull_bar(acc), acc_phase, te);
            else
                epilogue_tile<BLOCK_N, 2, CTAS, 0>(p, tacc, lane, half, grow, ncol0, ncols, tfull_bar(acc), acc_phase, te);
            if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
        }
        if (tma_out(MODE) && lane == 0) ptx::tma_store_wait_all<0>();    // shared memory must outlive the stores
    }

    // ------------------------------------------- teardown -------------------------------------------
    ptx::tc_fence_before();
    if (CTAS == 2) ptx::cluster_sync();      // neither CTA may exit while its peer can still touch its smem / TMEM
    else __syncthreads();
    if (warp_idx == 1) {
        ptx::tc_fence_after();
        if (CTAS == 2) ptx::tmem_dealloc_pair(tmem_base, C::TMEM_COLS);
        else ptx::tmem_dealloc(tmem_base, C::TMEM_COLS);
    }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
// tensormap_encoder() / make_tmap() (2-D bf16, [box_rows][64] SWIZZLE_128B boxes): runtime.cu, declared in common.cuh

// fp32 output [rows][cols], row pitch ld elements; box = 32 rows x 32 columns (128 B), SWIZZLE_128B
int make_tmap_out_f32(CUtensorMap* m, const void* base, long long rows, long long cols, long long ld) {
    EncodeTiledFn enc = tensormap_encoder();
    UC2_REQUIRE(enc != nullptr, UC2_ERR_CUDA, "cuTensorMapEncodeTiled entry point not found");
    cuuint64_t gdim[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
    cuuint64_t gstr[1] = {static_cast<cuuint64_t>(ld) * 4};
    cuuint32_t box[2] = {32u, 32u};
    cuuint32_t estr[2] = {1u, 1u};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), gdim, gstr, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    UC2_REQUIRE(r == CUDA_SUCCESS, UC2_ERR_CUDA, "cuTensorMapEncodeTiled(out f32) failed (%d): rows=%lld cols=%lld ld=%lld",
                (int)r, rows, cols, ld);
    return UC2_OK;
}

// The (counter, finished workers) pair of the next launch: SCHED_SLOTS pairs per device, zeroed once, handed out in
// turn; a launch leaves its pair zeroed (sched_draw), and a pair comes around again SCHED_SLOTS launches later.
int* next_sched_pair() {
    static std::mutex mu;
    static int* base[64] = {};
    static unsigned next[64] = {};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
    std::lock_guard<std::mutex> lk(mu);
    if (!base[dev]) {
        int* q = nullptr;
        if (cudaMalloc(&q, SCHED_SLOTS * 2 * sizeof(int)) != cudaSuccess) return nullptr;
        if (cudaMemset(q, 0, SCHED_SLOTS * 2 * sizeof(int)) != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess) {
            cudaFree(q);
            return nullptr;
        }
        base[dev] = q;
    }
    return base[dev] + 2 * (next[dev]++ % SCHED_SLOTS);
}

// How many persistent workers (CTAs, or CTA pairs) the device can hold for one kernel instantiation.
template <int BLOCK_N, bool A_MN, bool B_MN, int CTAS, int MODE>
int worker_slots(cudaError_t* err) {
    using C = Cfg<BLOCK_N, CTAS, MODE>;
    static std::once_flag once;
    static cudaError_t attr_err = cudaSuccess;
    static int slots = 0;
    std::call_once(once, [&] {
        auto kern = gemm_bf16_kernel<BLOCK_N, A_MN, B_MN, CTAS, MODE>;
        attr_err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES);
        slots = num_sms();
        if (CTAS == 2 && attr_err == cudaSuccess) {
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(2 * (num_sms() / 2));
            cfg.blockDim = dim3(gemm_threads(MODE));
            cfg.dynamicSmemBytes = C::SMEM_BYTES;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension;
            at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
            cfg.attrs = at; cfg.numAttrs = 1;
            int n = 0;
            if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) == cudaSuccess && n > 0) slots = n;
            else { slots = num_sms() / 2; cudaGetLastError(); }
        }
    });
    *err = attr_err;
    return slots;
}

template <int BLOCK_N, bool A_MN, bool B_MN, int CTAS, int MODE = 0>
int launch(const uc2_gemm_args& a, const GemmParams& p_in, cudaStream_t stream) {
    using C = Cfg<BLOCK_N, CTAS, MODE>;
    static_assert(C::STAGES >= 3, "pipeline too shallow");
    static_assert(CTAS == 1 || C::B_ROWS % 64 == 0, "a CTA pair needs BLOCK_N >= 128");
    auto kern = gemm_bf16_kernel<BLOCK_N, A_MN, B_MN, CTAS, MODE>;
    cudaError_t attr_err;
    const int slots = worker_slots<BLOCK_N, A_MN, B_MN, CTAS, MODE>(&attr_err);
    UC2_REQUIRE(attr_err == cudaSuccess, UC2_ERR_CUDA, "cudaFuncSetAttribute(smem=%d): %s", C::SMEM_BYTES,
                cudaGetErrorString(attr_err));
    CUtensorMap ta, tb;
    int rc;
    if (!A_MN) rc = make_tmap(&ta, a.a, a.M, a.K, a.lda, BLOCK_M);
    else       rc = make_tmap(&ta, a.a, a.K, a.M, a.lda, BLOCK_K);
    if (rc) return rc;
    if (!B_MN) rc = make_tmap(&tb, a.b, a.N, a.K, a.ldb, C::B_ROWS);
    else       rc = make_tmap(&tb, a.b, a.K, a.N, a.ldb, BLOCK_K);
    if (rc) return rc;
    GemmParams p = p_in;
    CUtensorMap tb_tail = tb;
    CUtensorMap tout = tb;                    // only read by the TMA-store modes
    if (tma_out(MODE)) {
        rc = make_tmap_out_f32(&tout, a.out_f32, a.M, a.N, a.ld_f32);
        if (rc) return rc;
    }
    const int tiles_mn = p.num_m_blocks * p.num_n_blocks;
    p.tail_start = tiles_mn;
    p.tail_split = 1;
    if (CTAS == 2 && BLOCK_N == 256 && p.split_k == 1 && a.tail_split != 1) {
        // the partial last round of the persistent pairs: cut its tiles into column slices (see GemmParams)
        const int full = tiles_mn / slots * slots, rem = tiles_mn - full;
        if (full > 0 && rem > 0) {
            int s = 1;
            if (rem * 2 <= slots) s = 2;
            if (!B_MN && rem * 4 <= slots) s = 4;                 // MN-major B boxes are 64 columns wide: halves only
            if (s > 1) {
                p.tail_start = full;
                p.tail_split = s;
                if (!B_MN) {
                    rc = make_tmap(&tb_tail, a.b, a.N, a.K, a.ldb, C::B_ROWS / s);
                    if (rc) return rc;
                }
            }
        }
    }
    const int total = p.tail_split > 1 ? p.tail_start + (tiles_mn - p.tail_start) * p.tail_split : tiles_mn * p.split_k;
    const int workers = total < slots ? total : slots;
    p.sched = nullptr;
    if (gemm_sched_dynamic() && total > workers) {
        p.sched = next_sched_pair();
        UC2_REQUIRE(p.sched != nullptr, UC2_ERR_CUDA, "gemm: could not allocate the tile scheduler counters");
    }
    {
        ProfScope prof(stream, 0, 2.0 * a.M * a.N * a.K);
        launch_pdl(kern, dim3(CTAS * workers), dim3(gemm_threads(MODE)), C::SMEM_BYTES, stream, CTAS, ta, tb, tb_tail, tout, p);
    }
    return check_last("gemm_bf16_kernel");
}

template <int BLOCK_N, int CTAS>
int dispatch_major(const uc2_gemm_args& a, const GemmParams& p, cudaStream_t s) {
    if (!a.a_mn && !a.b_mn) return launch<BLOCK_N, false, false, CTAS>(a, p, s);
    if (!a.a_mn && a.b_mn) return launch<BLOCK_N, false, true, CTAS>(a, p, s);
    if (a.a_mn && a.b_mn) return launch<BLOCK_N, true, true, CTAS>(a, p, s);
    return launch<BLOCK_N, true, false, CTAS>(a, p, s);
}

// Tile shape: fewest rounds of the persistent workers, weighted by the tile's MMA time; CTA pairs are preferred
// (a 128-row CTA re-reads B from L2 twice as often) whenever the problem has at least 256 rows.
void pick_tile(int M, int N, int split_k, int want_bn, int want_ctas, int* bn_out, int* ctas_out, bool splittable = false) {
    const int sms = num_sms();
    if (splittable && !want_bn && !want_ctas && M >= 2 * BLOCK_M && N >= 256) {
        // K gets split to fill the machine afterwards: take the tile with the best bytes-per-flop
        *bn_out = 256;
        *ctas_out = 2;
        return;
    }
    double best = 1e30;
    int best_bn = 128, best_ctas = 1;
    const int bns[3] = {256, 128, 64};
    for (int ctas = 2; ctas >= 1; --ctas) {
        if (want_ctas && ctas != want_ctas) continue;
        if (ctas == 2 && M <= BLOCK_M && !want_ctas) continue;
        for (int i = 0; i < 3; ++i) {
            const int bn = bns[i];
            if (want_bn && bn != want_bn) continue;
            if (ctas == 2 && bn < 128) continue;
            const int units = ctas == 2 ? sms / 2 : sms;
            const long long tiles = 1LL * ((M + BLOCK_M * ctas - 1) / (BLOCK_M * ctas)) * ((N + bn - 1) / bn) * split_k;
            const long long rounds = (tiles + units - 1) / units;
            // per-tile cost ~ MMA time (prop. to bn) + fixed overhead; narrow tiles and 128-row CTAs pay for their
            // extra shared-memory / L2 traffic per flop (measured: scripts/gemm_bench.py, profiles/)
            const double cost = rounds * (bn + 24.0) * (ctas == 1 ? 1.15 : 1.0) * (bn == 256 ? 1.0 : 1.3);
            if (cost < best - 1e-9) { best = cost; best_bn = bn; best_ctas = ctas; }
        }
    }
    *bn_out = best_bn;
    *ctas_out = best_ctas;
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
    UC2_REQUIRE(a.block_n == 0 || a.block_n == 64 || a.block_n == 128 || a.block_n == 256, UC2_ERR_ARG,
                "uc2_gemm_bf16: block_n must be 0/64/128/256");
    UC2_REQUIRE(a.ctas >= 0 && a.ctas <= 2 && !(a.ctas == 2 && a.block_n == 64), UC2_ERR_ARG,
                "uc2_gemm_bf16: ctas must be 0 (auto), 1 or 2 (2 needs block_n >= 128)");
    int split_k = a.split_k < 0 ? 1 : a.split_k;
    const bool can_split = a.out_f32 && a.accumulate && !a.out_bf16 && !a.out_pre && !a.bias && !a.residual &&
                           a.act == UC2_ACT_NONE;
    int bn = 0, ctas = 0;
    if (split_k == 0) {
        // auto: only meaningful for the atomic-accumulate (wgrad) form; fill the workers with K splits
        split_k = 1;
        pick_tile(a.M, a.N, 1, a.block_n, a.ctas, &bn, &ctas, can_split);
        if (can_split) {
            const long long tiles = 1LL * ((a.M + BLOCK_M * ctas - 1) / (BLOCK_M * ctas)) * ((a.N + bn - 1) / bn);
            const int kblocks = (a.K + BLOCK_K - 1) / BLOCK_K;
            const int units = ctas == 2 ? num_sms() / 2 : num_sms();
            double best = 0.0;
            for (int s = 1; s <= 32 && kblocks / s >= 4; ++s) {
                const long long t = tiles * s;
                const double eff = (double)t / (double)(((t + units - 1) / units) * units);
                if (eff > best + 0.03) { best = eff; split_k = s; }
            }
        }
    } else {
        pick_tile(a.M, a.N, split_k, a.block_n, a.ctas, &bn, &ctas);
    }
    if (split_k > 1)
        UC2_REQUIRE(can_split, UC2_ERR_ARG, "uc2_gemm_bf16: split_k>1 requires accumulate into out_f32 only");
    GemmParams p;
    p.M = a.M; p.N = a.N; p.K = a.K;
    p.num_m_blocks = (a.M + BLOCK_M * ctas - 1) / (BLOCK_M * ctas);
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
    p.drop.key = a.drop_key; p.drop.thresh = a.drop_thresh; p.drop.scale = a.drop_scale;
    p.ce_stats = reinterpret_cast<float2*>(a.ce_stats); p.ce_ld = a.ld_ce; p.ce_labels = a.ce_labels; p.ce_tgt = a.ce_tgt;
    UC2_REQUIRE(a.drop_thresh < 65536u, UC2_ERR_ARG, "uc2_gemm_bf16: drop_thresh must be < 65536");
    // 256-bit row slices: 32-byte aligned bases, pitches that keep every row 32-byte aligned
    auto al32 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 31) == 0; };
    p.vec_ok = (!a.residual || (a.ld_res % (a.residual_f32 ? 8 : 16) == 0 && al32(a.residual))) &&
               (!a.aux || (a.ld_aux % 16 == 0 && al32(a.aux))) &&
               (!a.out_bf16 || (a.ld_out % 16 == 0 && al32(a.out_bf16))) &&
               (!a.out_pre || (a.ld_pre % 16 == 0 && al32(a.out_pre))) &&
               (!a.out_f32 || (a.ld_f32 % 8 == 0 && al32(a.out_f32))) && (!a.bias || aligned16(a.bias));
    p.num_n_blocks = (a.N + bn - 1) / bn;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (a.ce_stats) {
        // MLM decoder + cross-entropy statistics (MODE 9): bf16 logits = A B^T + bias, nothing else
        UC2_REQUIRE(a.ce_labels && a.ce_tgt && a.ld_ce >= a.M && (reinterpret_cast<uintptr_t>(a.ce_stats) & 7) == 0,
                    UC2_ERR_ARG, "uc2_gemm_bf16: ce_stats needs ce_labels, ce_tgt and ld_ce >= M");
        UC2_REQUIRE(!a.a_mn && !a.b_mn && a.bias && a.out_bf16 && !a.out_f32 && !a.out_pre && !a.residual && !a.aux &&
                        a.act == UC2_ACT_NONE && a.drop_thresh == 0 && split_k == 1 && p.vec_ok,
                    UC2_ERR_ARG, "uc2_gemm_bf16: ce_stats goes with out_bf16 = A B^T + bias only (32-byte aligned rows)");
        p.num_m_blocks = (a.M + BLOCK_M * (a.M > BLOCK_M ? 2 : 1) - 1) / (BLOCK_M * (a.M > BLOCK_M ? 2 : 1));
        p.num_n_blocks = (a.N + 255) / 256;
        if (a.M > BLOCK_M) return launch<256, false, false, 2, 9>(a, p, s);
        return launch<256, false, false, 1, 9>(a, p, s);
    }
    if (ctas == 2 && bn == 256 && p.vec_ok) {
        // the BertLayer stack's own epilogues, compiled without run-time switches (see Epi<MODE>)
        const bool bf = a.out_bf16 && !a.out_f32, f32o = a.out_f32 && !a.out_bf16;
        const bool kk = !a.a_mn && !a.b_mn, kn = !a.a_mn && a.b_mn, nn = a.a_mn && a.b_mn;
        const bool res32 = a.residual && a.residual_f32, resb = a.residual && !a.residual_f32;
        if (a.drop_thresh != 0) {
            if (kk && a.bias && !a.out_pre && a.act == UC2_ACT_NONE && res32 && f32o && !a.accumulate)
                return launch<256, false, false, 2, 8>(a, p, s);
            return dispatch_major<256, 2>(a, p, s);
        }
        if (kk && a.bias && a.out_pre && a.act == UC2_ACT_GELU && !a.residual && bf)
            return launch<256, false, false, 2, 1>(a, p, s);
        if (kn && !a.bias && !a.out_pre && a.act == UC2_ACT_DGELU && !a.residual && bf)
            return launch<256, false, true, 2, 2>(a, p, s);
        if (kk && a.bias && !a.out_pre && a.act == UC2_ACT_NONE && res32 && f32o && !a.accumulate)
            return launch<256, false, false, 2, 3>(a, p, s);
        if (kk && a.bias && !a.out_pre && a.act == UC2_ACT_NONE && !a.residual && bf)
            return launch<256, false, false, 2, 4>(a, p, s);
        if (kn && !a.bias && !a.out_pre && a.act == UC2_ACT_NONE && resb && bf)
            return launch<256, false, true, 2, 5>(a, p, s);
        if (nn && !a.bias && !a.out_pre && a.act == UC2_ACT_NONE && !a.residual && f32o && a.accumulate)
            return launch<256, true, true, 2, 6>(a, p, s);
        if (kn && !a.bias && !a.out_pre && a.act == UC2_ACT_NONE && !a.residual && bf)
            return launch<256, false, true, 2, 7>(a, p, s);
    }
    if (ctas == 2) {
        if (bn == 256) return dispatch_major<256, 2>(a, p, s);
        return dispatch_major<128, 2>(a, p, s);
    }
    if (bn == 256) return dispatch_major<256, 1>(a, p, s);
    if (bn == 128) return dispatch_major<128, 1>(a, p, s);
    return dispatch_major<64, 1>(a, p, s);
}
