// tcgen05 / TMEM attention, forward (S <= 256) and backward (S <= 160), 12 heads x 64: the default kernels behind
// uc2_attention_fwd* / _bwd* for those lengths.  Replaces the same reference code as attention.cu
// (BertSelfAttention.forward model/layer.py:80-100) with every product on the 5th-generation tensor cores:
// accumulators in tensor memory, operands by TMA (SWIZZLE_128B), the forward's P kept in tensor memory as the TMEM-A
// operand of P V.  Each kernel's own comment block describes its pipeline.  First run on hardware in round 2; parity
// against a torch fp32 restatement, against the mma.sync kernels and through the whole model is in
// tests/test_attention_tc_gpu.py, tests/test_attention_gpu.py and tests/test_model_gpu.py; profiles/r02_* hold the
// ncu captures.  UC2_ATTN_PROF=1 prints per-warp phase cycle counters (debug only; it synchronises).
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
constexpr int TC_MAX_SP = 256;               // forward; the backward kernel below stops at TC_BWD_MAX_SP
constexpr int FWD_NSPLIT = 2;                // softmax warps per TMEM lane quarter (each takes a column range of the row)
constexpr int FWD_EW = 4 * FWD_NSPLIT;
constexpr int FWD_THREADS = 64 + 32 * FWD_EW;   // TMA warp, MMA warp, the softmax warps
constexpr float LOG2E = 1.4426950408889634f;
constexpr float SCALE_LOG2 = 0.125f * LOG2E;
constexpr float MASK_LOG2 = -10000.0f * LOG2E;
constexpr uint32_t P_CHUNK_BYTES = 128 * 128;   // 128 rows x 64 columns of bf16 (the backward's K-major operand tiles)

struct TcParams {
    const long long* mask;
    bf16* ctx;
    float* lse;
    int B, S, SP, items;
    DropCfg drop;
    long long* prof;      // debug (UC2_ATTN_PROF=1): per CTA, per warp, 8 cycle counters; nullptr otherwise
};
constexpr int PROF_SLOTS = 8;
struct PhaseClock {       // accumulates clock64 deltas of one warp's phases when profiling is on
    long long* out;
    long long t;
    __device__ __forceinline__ PhaseClock(long long* base, int warps_per_cta) {
        out = base ? base + ((long long)blockIdx.x * warps_per_cta + (threadIdx.x >> 5)) * PROF_SLOTS : nullptr;
        t = out ? clock64() : 0;
    }
    __device__ __forceinline__ void lap(int slot) {
        if (out) {
            const long long now = clock64();
            if ((threadIdx.x & 31) == 0) out[slot] += now - t;
            t = now;
        }
    }
};

// shared-memory / tensor-memory plan shared by host and device
struct TcSmem {
    uint32_t tile_bytes, off_q1, off_kv, off_mbias, off_info, off_red, off_bar, total;
    uint32_t tm_cols, tm_o;
    int nt, rem, rot_step, rot_n;
};
__host__ __device__ inline TcSmem tc_smem(int S, int SP) {
    TcSmem L;
    L.nt = S > 128 ? 2 : 1;
    L.tile_bytes = SP * 128u;
    // SP <= 160: 256 columns per CTA (S [0,160), O [160,224)) and two CTAs per SM; above: all 512 columns (S [0,256),
    // O [256,320)) and the SM to itself.  P overwrites S in place: a softmax warp owns the column range [c0, c1) of
    // its rows and stores the bf16 pairs of keys c.. at columns c0 + (c - c0) / 2, which it has already read.
    const bool small = SP <= 160;
    L.tm_cols = small ? 256u : 512u;
    L.tm_o = small ? 160u : 256u;
    // [Q rows 0..127 (16 KB)] [Q rows 128..SP-1] [K0 V0] [K1 V1]: Q single buffered (dead once the item's S products
    // retire), K / V double buffered.  The M = 128 reads of a Q tile cover 16 KB whatever SP is; what lies past the
    // valid rows is other tiles' data and lands in accumulator lanes nobody reads.
    // The remainder tile (rows 128..) holds `rem` valid rows.  A warp can only read the TMEM lane quarter warp % 4,
    // which is also its scheduler, so with the remainder always in lanes 0.. one scheduler would do twice the
    // softmax work of the others: the tile's base address is shifted down by rot rows instead (rot = rot_step *
    // ((item counter + CTA) % rot_n)), which moves the valid rows to lanes rot.. and spreads that work over time.
    L.rem = L.nt == 2 ? SP - 128 : 0;
    L.rot_step = L.rem <= 32 ? 32 : 64;
    L.rot_n = L.nt == 2 ? (L.rem <= 32 ? 4 : (L.rem <= 64 ? 2 : 1)) : 1;
    L.off_q1 = 16384u;
    L.off_kv = L.off_q1 + L.rem * 128u;
    L.off_mbias = L.off_kv + 4u * L.tile_bytes;
    L.off_info = L.off_mbias + 2u * SP * 4u;
    L.off_red = L.off_info + 32u;
    L.off_bar = L.off_red + FWD_EW * 32u * 8u;
    L.total = 1024u + L.off_bar + 128u;
    return L;
}

__host__ __device__ inline int active_warps(int S, int t) {
    const int rows = S - 128 * t;
    return rows <= 0 ? 0 : (rows >= 128 ? 4 : (rows + 31) >> 5);
}

// Pass 1 over 16 score columns of this thread's row.  PREFIX: every one of the 16 keys is valid (bias 0), the
// maximum is taken on the raw scores and scaled once per row; else the additive bias row is applied per element.
template <bool PREFIX>
__device__ __forceinline__ float max_chunk16(const uint32_t* r, const float* mb, float m) {
    if (PREFIX) {
#pragma unroll
        for (int j = 0; j < 16; j += 2) m = fmaxf(m, fmaxf(__uint_as_float(r[j]), __uint_as_float(r[j + 1])));
    } else {
#pragma unroll
        for (int j = 0; j < 16; j += 4) {
            const float4 bb = *reinterpret_cast<const float4*>(mb + j);
            m = fmaxf(m, fmaxf(fmaxf(fmaf(__uint_as_float(r[j]), SCALE_LOG2, bb.x),
                                     fmaf(__uint_as_float(r[j + 1]), SCALE_LOG2, bb.y)),
                               fmaxf(fmaf(__uint_as_float(r[j + 2]), SCALE_LOG2, bb.z),
                                     fmaf(__uint_as_float(r[j + 3]), SCALE_LOG2, bb.w))));
        }
    }
    return m;
}

// Pass 2 over 16 score columns: probabilities (relative to the row maximum m, exp2 domain), their sum into l, the
// attention-probability dropout of layer.py:94 (the normaliser keeps every key; the 1 / (1 - p) rescale is applied
// once per output element in the epilogue), bf16 pairs into 8 columns of the P tile in tensor memory.
template <bool PREFIX, bool DROP>
__device__ __forceinline__ void softmax_chunk16(const uint32_t* r, const float* mb, float m, float& l, uint32_t t32,
                                                uint32_t rbase, uint32_t tP) {
    uint32_t pk[8];
#pragma unroll
    for (int j = 0; j < 16; j += 4) {
        float p0, p1, p2, p3;
        if (PREFIX) {
            p0 = fast_ex2(fmaf(__uint_as_float(r[j]), SCALE_LOG2, -m));
            p1 = fast_ex2(fmaf(__uint_as_float(r[j + 1]), SCALE_LOG2, -m));
            p2 = fast_ex2(fmaf(__uint_as_float(r[j + 2]), SCALE_LOG2, -m));
            p3 = fast_ex2(fmaf(__uint_as_float(r[j + 3]), SCALE_LOG2, -m));
        } else {
            const float4 bb = *reinterpret_cast<const float4*>(mb + j);
            p0 = fast_ex2(fmaf(__uint_as_float(r[j]), SCALE_LOG2, bb.x) - m);
            p1 = fast_ex2(fmaf(__uint_as_float(r[j + 1]), SCALE_LOG2, bb.y) - m);
            p2 = fast_ex2(fmaf(__uint_as_float(r[j + 2]), SCALE_LOG2, bb.z) - m);
            p3 = fast_ex2(fmaf(__uint_as_float(r[j + 3]), SCALE_LOG2, bb.w) - m);
        }
        l += (p0 + p1) + (p2 + p3);
        if (DROP) {      // common.cuh "Attention-probability dropout": rbase = block hash * CA^(query & 15)
            p0 = rbase * drop_pow(DROP_CB, j) >= t32 ? p0 : 0.f;
            p1 = rbase * drop_pow(DROP_CB, j + 1) >= t32 ? p1 : 0.f;
            p2 = rbase * drop_pow(DROP_CB, j + 2) >= t32 ? p2 : 0.f;
            p3 = rbase * drop_pow(DROP_CB, j + 3) >= t32 ? p3 : 0.f;
        }
        pk[j / 2] = pack_bf16(p0, p1);
        pk[j / 2 + 1] = pack_bf16(p2, p3);
    }
    ptx::tmem_st_32x8(tP, pk);
}

__device__ __forceinline__ void quarter_sync(int q) {      // the FWD_NSPLIT warps that share TMEM lane quarter q
    asm volatile("bar.sync %0, %1;" ::"r"(1 + q), "n"(32 * FWD_NSPLIT) : "memory");
}

// One query row's share (columns [c0, c1), multiples of 16) of the two softmax passes.  The TMEM loads are software
// pipelined: the next 16 columns are in flight while the current 16 are processed (tcgen05.wait::ld waits for all).
// Columns [0, nfast) are chunks of valid keys only (prefix masks: all of them), the rest go through the bias row.
template <bool DROP>
__device__ __forceinline__ void softmax_row(uint32_t tS, const float* mb, int c0, int c1, int nfast,
                                            float* red_max, float* red_sum, int slot, int q, int lane, uint32_t t32,
                                            uint32_t hkey, uint32_t row, float& m_out, float& l_out) {
    // dropout: per-thread factor of the query row; the block hash is taken per 16-column chunk
    const uint32_t apow = DROP ? drop_pow_rt(DROP_CA, row) : 0u, iblk = row >> 4;
    auto rbase = [&](int c) { return DROP ? drop_block_hash(hkey, iblk, static_cast<uint32_t>(c) >> 4) * apow : 0u; };
    const uint32_t tP = tS + c0 - (c0 >> 1);            // + (c >> 1) = c0 + (c - c0) / 2: P in place, behind the reads
    uint32_t ra[16], rb[16];
    float m_raw = -INFINITY, m = -INFINITY, l = 0.f;
    // ---- pass 1
    if (c0 < c1) {
        ptx::tmem_ld_32x16(tS + c0, ra);
        ptx::tmem_wait_ld16(ra);
        for (int c = c0; c < c1; c += 32) {
            if (c + 16 < c1) ptx::tmem_ld_32x16(tS + c + 16, rb);
            if (c + 16 <= nfast) m_raw = max_chunk16<true>(ra, nullptr, m_raw);
            else m = max_chunk16<false>(ra, mb + c, m);
            ptx::tmem_wait_ld16(rb);
            if (c + 16 < c1) {
                if (c + 32 < c1) ptx::tmem_ld_32x16(tS + c + 32, ra);
                if (c + 32 <= nfast) m_raw = max_chunk16<true>(rb, nullptr, m_raw);
                else m = max_chunk16<false>(rb, mb + c + 16, m);
                ptx::tmem_wait_ld16(ra);
            }
        }
        ptx::tmem_ld_32x16(tS + c0, ra);                  // first chunk of pass 2, in flight across the exchange
    }
    m = fmaxf(m, m_raw * SCALE_LOG2);
    red_max[slot] = m;
    quarter_sync(q);
#pragma unroll
    for (int o = 0; o < FWD_NSPLIT; ++o) m = fmaxf(m, red_max[(o * 4 + q) * 32 + lane]);
    // ---- pass 2
    if (c0 < c1) {
        ptx::tmem_wait_ld16(ra);
        for (int c = c0; c < c1; c += 32) {
            if (c + 16 < c1) ptx::tmem_ld_32x16(tS + c + 16, rb);
            if (c + 16 <= nfast) softmax_chunk16<true, DROP>(ra, nullptr, m, l, t32, rbase(c), tP + (c >> 1));
            else softmax_chunk16<false, DROP>(ra, mb + c, m, l, t32, rbase(c), tP + (c >> 1));
            ptx::tmem_wait_ld16(rb);
            if (c + 16 < c1) {
                if (c + 32 < c1) ptx::tmem_ld_32x16(tS + c + 32, ra);
                if (c + 32 <= nfast)
                    softmax_chunk16<true, DROP>(rb, nullptr, m, l, t32, rbase(c + 16), tP + ((c + 16) >> 1));
                else
                    softmax_chunk16<false, DROP>(rb, mb + c + 16, m, l, t32, rbase(c + 16), tP + ((c + 16) >> 1));
                ptx::tmem_wait_ld16(ra);
            }
        }
    }
    red_sum[slot] = l;
    ptx::tmem_wait_st();
    quarter_sync(q);
    l = 0.f;
#pragma unroll
    for (int o = 0; o < FWD_NSPLIT; ++o) l += red_sum[(o * 4 + q) * 32 + lane];
    m_out = m;
    l_out = l;
}

// Forward.  CTAs loop over (batch, head) items; two CTAs share an SM for SP <= 160, so one CTA's tensor work and
// barrier latencies run under the other's softmax.  Per item and query tile t (128 rows; t = 1 holds rows 128..SP-1):
//     warp 0      TMA: K, V (double buffered) and Q (single buffered: it is dead once the item's S products retire)
//                 of the head as [SP x 64] SWIZZLE_128B boxes of the packed qkv tensor; the additive key-mask row in
//                 the exp2 domain and (kc, nfast): when the mask row is a prefix of ones -- every UC2 collate makes
//                 such rows -- only the first kc = ceil16(valid keys) columns are computed at all (a masked key's
//                 probability exp(s - 10000 - max) is exactly 0 in fp32 next to any valid key) and the chunks that
//                 hold valid keys only skip the bias arithmetic; any other 0/1 row takes the general path
//     warp 1      one thread issues  S_t = Q_t K^T  (M 128, N kc, K 64) into TMEM and, once the softmax warps have
//                 published P_t,  O_t = P_t V  with A = P_t read from TENSOR MEMORY (K-major) and B = V, MN-major,
//                 straight from its [key][d] rows
//     warps 2-9   a thread owns ONE query row (TMEM lane) and, with FWD_NSPLIT warps per lane quarter, a column range
//                 of it: row max / row sum are exchanged through shared memory between the parts, two passes over
//                 the scores in TMEM (max; exp2 + sum + the counter-hash dropout of attention.cu), P stored as bf16
//                 pairs back into tensor memory; then the epilogue O / rowsum -> ctx, lse.
template <bool DROP>
__global__ void __launch_bounds__(FWD_THREADS, 2)
attention_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmap_qkv, const __grid_constant__ CUtensorMap tmap_q0,
                        const __grid_constant__ CUtensorMap tmap_q1, const TcParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* gen = smem_raw + (base - ptx::smem_u32(smem_raw));
    const int S = p.S, SP = p.SP;
    const TcSmem L = tc_smem(S, SP);
    float* mbias = reinterpret_cast<float*>(gen + L.off_mbias);
    volatile int* info = reinterpret_cast<volatile int*>(gen + L.off_info);     // [buf][kc, nfast]
    float* red_max = reinterpret_cast<float*>(gen + L.off_red);
    float* red_sum = red_max + FWD_EW * 32;
    const uint32_t bar = base + L.off_bar;
    // barriers (8 B each): kv_full[2] kv_empty[2] q_full[2] q_empty[2] (per query tile) s_full p_full o_full o_empty,
    // then the TMEM base address
    auto kvfull_bar = [&](int b) { return bar + 8u * b; };
    auto kvempty_bar = [&](int b) { return bar + 16u + 8u * b; };
    auto qfull_bar = [&](int t) { return bar + 32u + 8u * t; };
    auto qempty_bar = [&](int t) { return bar + 48u + 8u * t; };
    const uint32_t sfull_bar = bar + 64u, pfull_bar = bar + 72u, ofull_bar = bar + 80u, oempty_bar = bar + 88u;
    const uint32_t tmem_ptr_addr = bar + 96u;
    volatile uint32_t* tmem_ptr_gen = reinterpret_cast<volatile uint32_t*>(gen + L.off_bar + 96);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        ptx::prefetch_tmap(&tmap_qkv);
        ptx::prefetch_tmap(&tmap_q0);
        ptx::prefetch_tmap(&tmap_q1);
        for (int b = 0; b < 2; ++b) {
            ptx::mbar_init(kvfull_bar(b), 2);          // TMA transaction arrive + mask-row arrive
            ptx::mbar_init(kvempty_bar(b), 1);
        }
        for (int t = 0; t < 2; ++t) {
            ptx::mbar_init(qfull_bar(t), 1);
            ptx::mbar_init(qempty_bar(t), 1);
        }
        ptx::mbar_init(sfull_bar, 1);
        ptx::mbar_init(pfull_bar, FWD_EW);             // every softmax warp arrives, also those without rows
        ptx::mbar_init(ofull_bar, 1);
        ptx::mbar_init(oempty_bar, FWD_EW);
        ptx::fence_barrier_init();
        ptx::fence_proxy_async();
    }
    if (warp == 1) {
        ptx::tmem_alloc(tmem_ptr_addr, L.tm_cols);
        ptx::tmem_relinquish();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_gen;
    griddep_sync();        // everything above may run under the tail of the QKV GEMM

    if (warp == 0) {
        // ===================================== producer =====================================
        PhaseClock pc(p.prof, FWD_THREADS / 32);
        int it = 0;
        for (int item = blockIdx.x; item < p.items; item += gridDim.x, ++it) {
            const int buf = it & 1;
            const uint32_t ph = (it >> 1) & 1u;
            const int b = item / NH, h = item - b * NH;
            pc.lap(0);
            // the mask row first (global loads in flight under the wait below): additive key mask of model.py:433-436
            // in the exp2 domain; padding columns [S, SP) never contribute
            const long long* mrow = p.mask + (long long)b * S;
            bool on[TC_MAX_SP / 32];
#pragma unroll
            for (int k = 0; k < TC_MAX_SP / 32; ++k) {
                const int j = lane + 32 * k;
                on[k] = j < S && mrow[j] != 0;
            }
            if (lane == 0) ptx::mbar_wait(kvempty_bar(buf), ph ^ 1u);   // the P V products of item it - 2 retired
            __syncwarp();
            pc.lap(1);
            if (lane == 0) {
                const uint32_t dst = base + L.off_kv + buf * 2u * L.tile_bytes;
                ptx::mbar_arrive_expect_tx(kvfull_bar(buf), 2u * L.tile_bytes);
                ptx::tma_load_2d(dst, &tmap_qkv, kvfull_bar(buf), HID + h * HD, b * S);
                ptx::tma_load_2d(dst + L.tile_bytes, &tmap_qkv, kvfull_bar(buf), 2 * HID + h * HD, b * S);
            }
            float* mb = mbias + buf * SP;
            int ones = 0, last = -1;
#pragma unroll
            for (int k = 0; k < TC_MAX_SP / 32; ++k) {
                const int j = lane + 32 * k;
                if (j < SP) mb[j] = j < S ? (on[k] ? 0.f : MASK_LOG2) : -INFINITY;
                if (on[k]) { ++ones; last = j; }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                ones += __shfl_xor_sync(0xffffffffu, ones, o);
                last = max(last, __shfl_xor_sync(0xffffffffu, last, o));
            }
            if (lane == 0) {
                const bool prefix = ones > 0 && last + 1 == ones;
                info[2 * buf] = prefix ? (ones + 15) & ~15 : SP;      // kc: score columns computed
                info[2 * buf + 1] = prefix ? ones & ~15 : 0;          // nfast: leading columns without bias arithmetic
            }
            __syncwarp();
            pc.lap(2);
            if (lane == 0) {
                ptx::mbar_arrive(kvfull_bar(buf));
                // a query tile's buffer is free once the previous item's S product of that tile retired
                ptx::mbar_wait(qempty_bar(0), (it & 1u) ^ 1u);
                ptx::mbar_arrive_expect_tx(qfull_bar(0), (SP < 128 ? SP : 128) * 128u);
                ptx::tma_load_2d(base, &tmap_q0, qfull_bar(0), h * HD, b * S);
                if (L.nt == 2) {
                    ptx::mbar_wait(qempty_bar(1), (it & 1u) ^ 1u);
                    ptx::mbar_arrive_expect_tx(qfull_bar(1), L.rem * 128u);
                    ptx::tma_load_2d(base + L.off_q1, &tmap_q1, qfull_bar(1), h * HD, b * S + 128);
                }
            }
            __syncwarp();
            pc.lap(3);
        }
    } else if (warp == 1) {
        // ===================================== MMA issuer ===================================
        // Issue order: S(0); then per unit n: [P_n published] O_n = P_n V, and straight behind it S(n+1) -- the tensor
        // pipe runs in issue order, so S(n+1) overwrites the S / P columns only after O_n has read P_n, and it is
        // ready by the time the softmax warps are through with the epilogue of unit n.
        if (lane == 0) {
            const uint32_t idesc_o = ptx::idesc_bf16_f32(128, HD, false, true);
            const int my_items = p.items > (int)blockIdx.x ? (p.items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
            const uint32_t units = static_cast<uint32_t>(my_items) * L.nt;
            PhaseClock pc(p.prof, FWD_THREADS / 32);
            auto issue_s = [&](uint32_t n) {
                const int it = n / L.nt, t = n - it * L.nt, buf = it & 1;
                pc.lap(0);
                if (t == 0) ptx::mbar_wait(kvfull_bar(buf), (it >> 1) & 1u);
                ptx::mbar_wait(qfull_bar(t), it & 1u);
                pc.lap(1);
                ptx::tc_fence_after();
                const int kc = info[2 * buf];
                const uint32_t idesc_s = ptx::idesc_bf16_f32(128, kc, false, false);
                const uint32_t sK = base + L.off_kv + buf * 2u * L.tile_bytes;
                const uint32_t rot = L.rot_step * ((it + blockIdx.x) % L.rot_n);
                // Tile 1: the valid rows sit at M rows rot.. of a 128-row window over whatever surrounds them in
                // shared memory (those results are never read).
                const uint32_t sQ = t == 0 ? base : base + L.off_q1 - rot * 128u;
                // one descriptor per operand, advanced by 32 bytes (2 in the 16-byte address field) per 16 of head dim
                const uint64_t da = ptx::smem_desc_sw128(sQ, 16u, 1024u), db = ptx::smem_desc_sw128(sK, 16u, 1024u);
#pragma unroll
                for (int k = 0; k < HD / 16; ++k) ptx::umma_bf16(tmem_base, da + 2u * k, db + 2u * k, idesc_s, k > 0 ? 1u : 0u);
                ptx::umma_commit(sfull_bar);
                ptx::umma_commit(qempty_bar(t));           // this query tile is dead once the product retires
                pc.lap(2);
            };
            if (units > 0) issue_s(0);
            for (uint32_t n = 0; n < units; ++n) {
                const int it = n / L.nt, t = n - it * L.nt, buf = it & 1;
                const int kc = info[2 * buf];
                const int nch = kc >> 4, per = (nch + FWD_NSPLIT - 1) / FWD_NSPLIT;
                const uint32_t sV = base + L.off_kv + buf * 2u * L.tile_bytes + L.tile_bytes;
                pc.lap(0);
                ptx::mbar_wait(pfull_bar, n & 1u);
                pc.lap(3);
                if (n > 0) ptx::mbar_wait(oempty_bar, (n - 1u) & 1u);      // the previous unit's O has been read out
                pc.lap(4);
                ptx::tc_fence_after();
                // O = P V: A = the bf16 P tile in tensor memory (8 columns per 16 keys, each softmax part's share at the
                // start of its own column range), B = V rows, MN-major
                uint64_t dv = ptx::smem_desc_sw128(sV, 8192u, 1024u);      // + 2048 bytes (128) per 16 keys
                uint32_t acc = 0u;
                for (int part = 0; part < FWD_NSPLIT; ++part) {
                    uint32_t a = tmem_base + part * per * 16;
                    const int cnt = min(per, nch - part * per);
                    for (int i = 0; i < cnt; ++i, a += 8u, dv += 128u, acc = 1u)
                        ptx::umma_bf16_ts(tmem_base + L.tm_o, a, dv, idesc_o, acc);
                }
                ptx::umma_commit(ofull_bar);
                if (t == L.nt - 1) ptx::umma_commit(kvempty_bar(buf));     // K, V of this buffer are dead once these retire
                pc.lap(5);
                if (n + 1 < units) issue_s(n + 1);
            }
        }
    } else {
        // ===================================== softmax warps ================================
        const int q = warp & 3;                     // TMEM lane quarter this warp may access
        const int part = (warp - 2) >> 2;           // which column range of the row
        const uint32_t lane_sel = static_cast<uint32_t>(q * 32) << 16;
        const uint32_t tS = tmem_base + lane_sel;
        const uint32_t tO = tmem_base + lane_sel + L.tm_o + part * (HD / FWD_NSPLIT);
        const int slot = (part * 4 + q) * 32 + lane;
        const float out_scale = DROP ? p.drop.scale : 1.f;
        // epilogue of a unit: O / rowsum -> ctx (heads merged: layer.py:98-100 is free), lse for the backward.  It is
        // deferred until the NEXT unit's scores have arrived: S(n+1) is issued behind O_n = P_n V and the tensor pipe
        // runs in order, so s_full(n+1) also says O_n is complete -- one barrier round trip per unit instead of two.
        struct Pending { bf16* orow; float* lse; float m, l; bool active, row_ok; } prev = {nullptr, nullptr, 0.f, 1.f, false, false};
        auto epilogue = [&](const Pending& e) {
            if (e.active) {
                uint32_t r[32];
                constexpr int W = HD / FWD_NSPLIT;
                static_assert(W == 32, "the epilogue reads one 32-column block per part");
                ptx::tmem_ld_32x32(tO, r);
                ptx::tmem_wait_ld();
                if (e.row_ok) {
                    const float inv = out_scale / e.l;
#pragma unroll
                    for (int g = 0; g < 2; ++g) {
                        uint32_t o[8];
#pragma unroll
                        for (int j = 0; j < 8; ++j)
                            o[j] = pack_bf16(__uint_as_float(r[16 * g + 2 * j]) * inv,
                                             __uint_as_float(r[16 * g + 2 * j + 1]) * inv);
                        ptx::stg256(e.orow + 16 * g, o[0], o[1], o[2], o[3], o[4], o[5], o[6], o[7]);
                    }
                    if (part == 0) *e.lse = (e.m + log2f(e.l)) * (1.f / LOG2E);
                }
            }
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(oempty_bar);
        };
        PhaseClock pc(p.prof, FWD_THREADS / 32);
        int it = 0;
        uint32_t n = 0;
        for (int item = blockIdx.x; item < p.items; item += gridDim.x, ++it) {
            const int buf = it & 1;
            const int b = item / NH, h = item - b * NH;
            const float* mb = mbias + buf * SP;
            const uint32_t hkey = drop_head_key(p.drop.key, b * NH + h);
            pc.lap(0);
            ptx::mbar_wait(kvfull_bar(buf), (it >> 1) & 1u);          // the mask row and (kc, nfast) are there
            pc.lap(1);
            const int kc = info[2 * buf], nfast = info[2 * buf + 1];
            const int nch = kc >> 4, per = (nch + FWD_NSPLIT - 1) / FWD_NSPLIT;
            const int c_begin = min(part * per, nch) * 16, c_end = min((part + 1) * per, nch) * 16;
            const int rot = L.rot_step * ((it + blockIdx.x) % L.rot_n);
            for (int t = 0; t < L.nt; ++t, ++n) {
                // tile 0: lanes = rows 0..127; tile 1: lanes rot.. = rows 128.. (see tc_smem)
                const int rel = q * 32 + lane - (t ? rot : 0);
                const int row = t * 128 + rel;                        // query row inside the (batch, head)
                Pending cur;
                cur.row_ok = rel >= 0 && row < S;
                cur.active = __any_sync(0xffffffffu, cur.row_ok);     // uniform over the warps of a lane quarter
                cur.orow = p.ctx + ((long long)b * S + row) * HID + h * HD + part * (HD / FWD_NSPLIT);
                cur.lse = p.lse + ((long long)b * NH + h) * S + row;
                cur.m = 0.f; cur.l = 1.f;
                pc.lap(0);
                ptx::mbar_wait(sfull_bar, n & 1u);
                pc.lap(2);
                ptx::tc_fence_after();
                if (n > 0) epilogue(prev);
                pc.lap(3);
                if (cur.active)
                    softmax_row<DROP>(tS, mb, c_begin, c_end, nfast, red_max, red_sum, slot, q, lane,
                                      p.drop.thresh << 16, hkey, static_cast<uint32_t>(cur.row_ok ? row : 0), cur.m, cur.l);
                ptx::tc_fence_before();          // S_t reads and P_t stores are complete before the barrier
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive(pfull_bar);
                pc.lap(cur.active ? 4 : 5);
                prev = cur;
            }
        }
        if (n > 0) {
            ptx::mbar_wait(ofull_bar, (n - 1u) & 1u);
            ptx::tc_fence_after();
            epilogue(prev);
        }
    }

    // ------------------------------------------- teardown -------------------------------------------
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc(tmem_base, L.tm_cols);
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
// Tiles of Q, K, V, dO and O come by TMA once per item (single buffered; the next item's tiles are prefetched into
// L2 meanwhile); delta is the row-wise dot product of the dO and O tiles (both carry the same swizzle).  16
// element-wise warps = 4 TMEM lane quarters x 4 column parts; the dropout mask is the forward's (16 x 16 block hash,
// common.cuh), here with the key as the per-thread factor and the query as the compile-time one.
// =================================================================================================
struct TcBwdParams {
    const long long* mask;
    const bf16* ctx;
    const bf16* dctx;
    const float* lse;
    bf16* dqkv;
    int B, S, SP, items;
    DropCfg drop;
    long long* prof;      // debug (UC2_ATTN_PROF=1), see TcParams
    int blocks;           // 0: query blocks chosen from SP; 2: two blocks also for 128 < SP <= 160 (tuning knob)
};

struct TcBwdSmem {
    uint32_t tile_bytes, pt_bytes, off_o, off_ds, off_pd, off_vec, off_bar, total;
    uint32_t tm_s, tm_dp, tm_dv, tm_dk, tm_dq;
    int nu, nv, nchunk, qb0;       // qb0: queries in block 0 when nv == 2 (block 1 holds the rest)
};
// Two shapes of the same pipeline:
//   SP <= 160  one query block (nv = 1): S^T_u and dP^T_u hold all SP query columns (TMEM [0,SP) and [192,192+SP)), dV_u / dK_u
//              reuse the S^T columns [0,128) once the element-wise warps are through, dQ lives in [384,512)
//   SP <= 256  two query blocks of <= 128 queries (nv = 2): S^T_uv [0,128), dP^T_uv [128,256), dV_u [256,320) and dK_u
//              [320,384) accumulate over the blocks, dQ [384,512); the O tile (only needed for delta at the start of an
//              item) shares its shared memory with the dS^T / Pd^T tiles, which hold one query block at a time
__host__ __device__ inline TcBwdSmem tc_bwd_smem(int S, int SP, int blocks = 0) {
    TcBwdSmem L;
    const bool wide = SP > 160 || (blocks == 2 && SP > 128);
    L.nu = S > 128 ? 2 : 1;
    L.nv = wide ? 2 : 1;
    L.qb0 = SP > 160 ? 128 : 96;                     // (tuning knob blocks == 2 at 128 < SP <= 160: 96 + the rest)
    L.nchunk = wide ? 2 : (SP + 63) >> 6;
    L.tile_bytes = SP * 128u;
    L.pt_bytes = L.nchunk * P_CHUNK_BYTES;
    L.off_ds = (wide ? 4u : 5u) * L.tile_bytes;      // [Q][K][V][dO]([O]) then dS^T, then Pd^T (so MN-major over-reads of
    L.off_pd = L.off_ds + L.pt_bytes;                // dS^T for the padding query blocks of dQ stay inside the allocation)
    L.off_o = wide ? L.off_ds : 4u * L.tile_bytes;
    L.off_vec = L.off_pd + L.pt_bytes;               // 2 buffers x (lse2[SP], delta[SP])
    L.off_bar = L.off_vec + 4u * SP * 4u;
    L.total = 1024u + L.off_bar + 128u;
    L.tm_s = 0u;
    L.tm_dp = wide ? 128u : 192u;
    L.tm_dv = wide ? 256u : 0u;
    L.tm_dk = L.tm_dv + 64u;
    L.tm_dq = 384u;
    return L;
}

constexpr int TC_BWD_MAX_SP = 256;
constexpr uint32_t TMEM_COLS = 512;

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
                        const __grid_constant__ CUtensorMap tmap_o, const TcBwdParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* gen = smem_raw + (base - ptx::smem_u32(smem_raw));
    const int S = p.S, SP = p.SP;
    const TcBwdSmem L = tc_bwd_smem(S, SP, p.blocks);
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
        ptx::prefetch_tmap(&tmap_o);
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
    const uint8_t* gdO = gen + 3u * L.tile_bytes;        // generic-address views of the dO and O tiles (delta)
    const uint8_t* gO = gen + L.off_o;
    const uint32_t sDS = base + L.off_ds, sPD = base + L.off_pd;
    const int nv = L.nv;
    auto nq_of = [&](int v) { return nv == 1 ? SP : (v == 0 ? L.qb0 : SP - L.qb0); };    // queries of block v
    const uint32_t qb_bytes = static_cast<uint32_t>(L.qb0) * 128u;                   // Q / dO rows of block 1 start here

    if (warp == 0) {
        // ===================================== producer =====================================
        if (lane == 0) {
            int it = 0;
            for (int item = blockIdx.x; item < p.items; item += gridDim.x, ++it) {
                const int b = item / NH, h = item - b * NH;
                ptx::mbar_wait(ld_empty, (it & 1u) ^ 1u);          // every MMA of the previous item retired
                ptx::mbar_arrive_expect_tx(ld_full, 5u * L.tile_bytes);
                ptx::tma_load_2d(sQ, &tmap_qkv, ld_full, h * HD, b * S);
                ptx::tma_load_2d(sK, &tmap_qkv, ld_full, HID + h * HD, b * S);
                ptx::tma_load_2d(sV, &tmap_qkv, ld_full, 2 * HID + h * HD, b * S);
                ptx::tma_load_2d(sdO, &tmap_do, ld_full, h * HD, b * S);
                ptx::tma_load_2d(base + L.off_o, &tmap_o, ld_full, h * HD, b * S);
                const int nxt = item + gridDim.x;                  // warm L2 with the next item's tiles
                if (nxt < p.items) {
                    const int nb = nxt / NH, nh = nxt - nb * NH;
                    ptx::tma_prefetch_2d(&tmap_qkv, nh * HD, nb * S);
                    ptx::tma_prefetch_2d(&tmap_qkv, HID + nh * HD, nb * S);
                    ptx::tma_prefetch_2d(&tmap_qkv, 2 * HID + nh * HD, nb * S);
                    ptx::tma_prefetch_2d(&tmap_do, nh * HD, nb * S);
                    ptx::tma_prefetch_2d(&tmap_o, nh * HD, nb * S);
                }
            }
        }
    } else if (warp == 1) {
        // ===================================== MMA issuer ===================================
        if (lane == 0) {
            const uint32_t idesc_kv = ptx::idesc_bf16_f32(128, HD, false, true);
            const uint32_t idesc_dq = ptx::idesc_bf16_f32(128, HD, true, true);
            PhaseClock pc(p.prof, bwd_threads(NSPLIT) / 32);
            int it = 0;
            for (int item = blockIdx.x; item < p.items; item += gridDim.x, ++it) {
                const uint32_t itp = it & 1u;
                pc.lap(0);
                ptx::mbar_wait(ld_full, itp);
                pc.lap(1);
                ptx::tc_fence_after();
                for (int u = 0; u < nu; ++u) {
                    // the dV / dK of the previous key tile have been read out (they share columns with S^T when nv == 1)
                    pc.lap(0);
                    if (u == 0) ptx::mbar_wait(s_empty(nu - 1), itp ^ 1u);
                    else ptx::mbar_wait(s_empty(u - 1), itp);
                    pc.lap(2);
                    ptx::tc_fence_after();
                    const int ksteps = (u == 0 ? (SP < 128 ? SP : 128) : SP - 128) / 16;       // keys of this tile / 16
                    for (int v = 0; v < nv; ++v) {
                        const uint32_t ph = static_cast<uint32_t>(it * nv + v) & 1u;
                        const int nq = nq_of(v) / 16;                                        // queries of this block / 16
                        const uint32_t idesc_s = ptx::idesc_bf16_f32(128, nq * 16, false, false);
                        // S^T_uv = K_u Q_v^T, dP^T_uv = V_u dO_v^T.  Key tile 1 reads 128 rows from row 128 of K / V on;
                        // rows past SP are the following tiles.  Products of independent accumulators are interleaved.
                        {
                            const uint64_t dk = ptx::smem_desc_sw128(sK + u * 16384u, 16u, 1024u);
                            const uint64_t dq = ptx::smem_desc_sw128(sQ + v * qb_bytes, 16u, 1024u);
                            const uint64_t dv = ptx::smem_desc_sw128(sV + u * 16384u, 16u, 1024u);
                            const uint64_t dd = ptx::smem_desc_sw128(sdO + v * qb_bytes, 16u, 1024u);
#pragma unroll
                            for (int k = 0; k < HD / 16; ++k) {      // + 32 bytes (2 in the address field) per 16 of head dim
                                ptx::umma_bf16(tmem_base + L.tm_s, dk + 2u * k, dq + 2u * k, idesc_s, k > 0 ? 1u : 0u);
                                ptx::umma_bf16(tmem_base + L.tm_dp, dv + 2u * k, dd + 2u * k, idesc_s, k > 0 ? 1u : 0u);
                            }
                        }
                        ptx::umma_commit(sd_full(u));
                        pc.lap(3);

                        ptx::mbar_wait(pds_full(u), ph);               // Pd^T_uv, dS^T_uv written; S^T, dP^T read
                        pc.lap(4);
                        if (u == 0 && v == 0) ptx::mbar_wait(dq_empty, itp ^ 1u);   // previous item's dQ read out
                        pc.lap(5);
                        ptx::tc_fence_after();
                        // dV_u (+)= Pd^T_uv dO_v, dK_u (+)= dS^T_uv Q_v (K dimension = the block's queries: A = the K-major
                        // tiles the element-wise warps wrote, B = dO / Q rows, MN-major) and dQ (+)= dS K_u for the query
                        // tiles the block covers (A = the dS^T tile read MN-major: 64-query blocks P_CHUNK_BYTES apart,
                        // 16 key rows per K step; B = K rows of tile u, MN-major): one K step of each per round
                        {
                            const uint64_t b_do = ptx::smem_desc_sw128(sdO + v * qb_bytes, 8192u, 1024u);
                            const uint64_t b_q = ptx::smem_desc_sw128(sQ + v * qb_bytes, 8192u, 1024u);
                            const uint64_t b_k = ptx::smem_desc_sw128(sK + u * 16384u, 8192u, 1024u);
                            const uint64_t a_dq0 = ptx::smem_desc_sw128(sDS, P_CHUNK_BYTES, 1024u);
                            const uint64_t a_dq1 = ptx::smem_desc_sw128(sDS + 2u * P_CHUNK_BYTES, P_CHUNK_BYTES, 1024u);
                            const bool two_dq = nv == 1 && nu == 2;       // one block holding both query tiles
                            const uint32_t dq_col = L.tm_dq + (nv == 2 ? 64u * v : 0u);
                            const int rounds = nq > ksteps ? nq : ksteps;
                            for (int kk = 0; kk < rounds; ++kk) {
                                if (kk < nq) {
                                    const uint32_t a_off = (kk >> 2) * P_CHUNK_BYTES + (kk & 3) * 32u;
                                    const uint32_t acc = (v > 0 || kk > 0) ? 1u : 0u;
                                    ptx::umma_bf16(tmem_base + L.tm_dv, ptx::smem_desc_sw128(sPD + a_off, 16u, 1024u),
                                                   b_do + 128u * kk, idesc_kv, acc);
                                    ptx::umma_bf16(tmem_base + L.tm_dk, ptx::smem_desc_sw128(sDS + a_off, 16u, 1024u),
                                                   b_q + 128u * kk, idesc_kv, acc);
                                }
                                if (kk < ksteps) {
                                    const uint32_t acc = (u > 0 || kk > 0) ? 1u : 0u;
                                    ptx::umma_bf16(tmem_base + dq_col, a_dq0 + 128u * kk, b_k + 128u * kk, idesc_dq, acc);
                                    if (two_dq)
                                        ptx::umma_bf16(tmem_base + L.tm_dq + 64u, a_dq1 + 128u * kk, b_k + 128u * kk, idesc_dq, acc);
                                }
                            }
                        }
                        pc.lap(6);
                    }
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
        PhaseClock pc(p.prof, bwd_threads(NSPLIT) / 32);
        int it = 0;
        for (int item = blockIdx.x; item < p.items; item += gridDim.x, ++it) {
            const uint32_t itp = it & 1u;
            const int b = item / NH, h = item - b * NH;
            const long long row0 = (long long)b * S;
            pc.lap(0);
            // ---- per-query vectors: lse in the exp2 domain (+inf for padding columns -> P = 0), delta
            float* lse2 = vec + (it & 1) * 2 * SP;
            float* delta = lse2 + SP;
            {
                const float* Lg = p.lse + ((long long)b * NH + h) * S;
                for (int i = threadIdx.x - 64; i < SP; i += 128 * NSPLIT) lse2[i] = i < S ? Lg[i] * LOG2E : INFINITY;
                // delta = rowsum(dO * O) from the two TMA tiles: both carry the same 128-byte swizzle, so the products of
                // matching 16-byte units sum to the row's dot product whatever the unit order is
                ptx::mbar_wait(ld_full, itp);
                pc.lap(1);
                for (int r0 = ew * 4; r0 < SP; r0 += 16 * NSPLIT) {
                    const int r = r0 + (lane >> 3), off = r * 128 + (lane & 7) * 16;
                    float acc = 0.f;
                    if (r < SP) {
                        const uint4 ua = *reinterpret_cast<const uint4*>(gO + off);
                        const uint4 ud = *reinterpret_cast<const uint4*>(gdO + off);
                        const uint32_t a[4] = {ua.x, ua.y, ua.z, ua.w}, d[4] = {ud.x, ud.y, ud.z, ud.w};
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const float2 fa = unpack_bf16(a[k]), fd = unpack_bf16(d[k]);
                            acc = fmaf(fa.x, fd.x, fmaf(fa.y, fd.y, acc));
                        }
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
                const uint32_t prow = static_cast<uint32_t>(kr) * 128u;
                // dropout (common.cuh "Attention-probability dropout"): this thread's key contributes CB^(key & 15)
                const uint32_t bpow = drop_pow_rt(DROP_CB, static_cast<uint32_t>(key)), jblk = static_cast<uint32_t>(key) >> 4;
                const uint32_t t32 = p.drop.thresh << 16;
                for (int v = 0; v < nv; ++v) {
                    const uint32_t ph = static_cast<uint32_t>(it * nv + v) & 1u;
                    const int q_base = v * L.qb0;                                      // first query of the block
                    const int nch = nq_of(v) / 16, per = (nch + NSPLIT - 1) / NSPLIT;   // 16-column chunks per part
                    const int c_begin = min(part * per, nch) * 16, c_end = min((part + 1) * per, nch) * 16;
                    pc.lap(0);
                    ptx::mbar_wait(sd_full(u), ph);
                    pc.lap(2);
                    ptx::tc_fence_after();
                    for (int c = c_begin; c < c_end; c += 16) {
                        uint32_t rs[16], rd[16], pd[8], ds[8];
                        ptx::tmem_ld_32x16(tmem_base + lane_sel + L.tm_s + c, rs);
                        ptx::tmem_ld_32x16(tmem_base + lane_sel + L.tm_dp + c, rd);
                        const uint32_t rbase =
                            p.drop.thresh ? drop_block_hash(hkey, static_cast<uint32_t>(q_base + c) >> 4, jblk) * bpow : 0u;
                        ptx::tmem_wait_ld();
#pragma unroll
                        for (int j = 0; j < 16; j += 4) {
                            const float4 l4 = *reinterpret_cast<const float4*>(lse2 + q_base + c + j);
                            const float4 d4 = *reinterpret_cast<const float4*>(delta + q_base + c + j);
                            // two queries per instruction (common.cuh: packed fp32 pairs, same roundings as the scalar
                            // form): these loops are bound by instruction issue on the scheduler that owns the lane quarter
                            const f32x2 nlq[2] = {pk2(-l4.x, -l4.y), pk2(-l4.z, -l4.w)};
                            const f32x2 ndq[2] = {pk2(-d4.x, -d4.y), pk2(-d4.z, -d4.w)};
#pragma unroll
                            for (int e = 0; e < 2; ++e) {
                                const int je = j + 2 * e;
                                const f32x2 x = add2(fma2(pk2u(rs[je], rs[je + 1]), splat2(SCALE_LOG2), splat2(bias)), nlq[e]);
                                float x0, x1;
                                up2(x, x0, x1);
                                const f32x2 pe = pk2(fast_ex2(x0), fast_ex2(x1));
                                f32x2 dp = pk2u(rd[je], rd[je + 1]);
                                f32x2 pdv = pe;
                                if (p.drop.thresh) {
                                    // elements (query q_base + c + je (+1), key): same stream as the forward
                                    const f32x2 sc = pk2(rbase * drop_pow(DROP_CA, je) >= t32 ? p.drop.scale : 0.f,
                                                         rbase * drop_pow(DROP_CA, je + 1) >= t32 ? p.drop.scale : 0.f);
                                    pdv = mul2(pe, sc);                                   // finite operands only
                                    dp = mul2(dp, sc);
                                }
                                // dS without the 1/sqrt(64): a power of two commutes with the bf16 rounding, so it is
                                // applied once per OUTPUT element in the dK / dQ epilogues instead of once per score here
                                const f32x2 dsv = mul2(pe, add2(dp, ndq[e]));
                                pd[j / 2 + e] = pack_bf16x2(pdv);
                                ds[j / 2 + e] = pack_bf16x2(dsv);
                            }
                        }
#pragma unroll
                        for (int g = 0; g < 2; ++g) {
                            const uint32_t q0 = c + 8 * g;         // 8 queries = one 16-byte unit of the swizzled row
                            const uint32_t off = (q0 >> 6) * P_CHUNK_BYTES + prow + ((((q0 & 63u) >> 3) ^ sw) << 4);
                            ptx::st_shared_v4(sPD + off, pd[4 * g], pd[4 * g + 1], pd[4 * g + 2], pd[4 * g + 3]);
                            ptx::st_shared_v4(sDS + off, ds[4 * g], ds[4 * g + 1], ds[4 * g + 2], ds[4 * g + 3]);
                        }
                    }
                    ptx::fence_proxy_async();
                    ptx::tc_fence_before();
                    __syncwarp();
                    if (lane == 0) ptx::mbar_arrive(pds_full(u));
                    pc.lap(3);
                }

                // this key row's [dV_u | dK_u] = 128 adjacent TMEM columns: part -> a 128 / NSPLIT-column slice -> dqkv
                ptx::mbar_wait(kv_full(u), itp);
                pc.lap(4);
                ptx::tc_fence_after();
                constexpr int WKV = 128 / NSPLIT;
                const int ckv = part * WKV;
                bf16* dst = p.dqkv + (row0 + key) * QKV_LD + (ckv < HD ? 2 * HID : HID) + h * HD + (ckv & (HD - 1));
                store_cols<WKV>(tmem_base + lane_sel + L.tm_dv + ckv, dst, key_ok, ckv < HD ? 1.f : 0.125f);
                ptx::tc_fence_before();
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive(s_empty(u));
                pc.lap(5);
            }

            // dQ: query rows in the lanes; part = which 64 / NSPLIT of the 64 columns
            if (q * 32 < S) {
                pc.lap(0);
                ptx::mbar_wait(dq_full, itp);
                pc.lap(6);
                ptx::tc_fence_after();
                constexpr int WQ = HD / NSPLIT;
                // dQ tile 0 holds the queries from 0 on, tile 1 those from q1 on (128, or the second query block's start)
                const int q1 = nv == 2 ? L.qb0 : 128;
                const int rows0 = nv == 2 ? L.qb0 : (S < 128 ? S : 128);
                store_cols<WQ>(tmem_base + lane_sel + L.tm_dq + part * WQ,
                               p.dqkv + (row0 + kr) * QKV_LD + h * HD + part * WQ, kr < rows0 && kr < S, 0.125f);
                if (q1 + q * 32 < S)                   // second query tile: rows q1 + (this lane quarter's rows)
                    store_cols<WQ>(tmem_base + lane_sel + L.tm_dq + 64u + part * WQ,
                                   p.dqkv + (row0 + q1 + kr) * QKV_LD + h * HD + part * WQ, q1 + kr < S, 0.125f);
                ptx::tc_fence_before();
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive(dq_empty);
                pc.lap(7);
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
        const char* e = getenv("UC2_ATTN_TCGEN05");          // on unless UC2_ATTN_TCGEN05=0
        v = (e && e[0] == '0') ? 0 : 1;
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
    const TcSmem L = tc_smem(S, SP);
    CUtensorMap tmap, tq0, tq1;
    if (int rc = make_tmap(&tmap, qkv, (long long)B * S, QKV_LD, QKV_LD, SP)) return rc;
    if (int rc = make_tmap(&tq0, qkv, (long long)B * S, QKV_LD, QKV_LD, SP < 128 ? SP : 128)) return rc;
    if (int rc = make_tmap(&tq1, qkv, (long long)B * S, QKV_LD, QKV_LD, L.rem > 0 ? L.rem : 16)) return rc;
    static std::once_flag once;
    static cudaError_t attr_err = cudaSuccess;
    std::call_once(once, [] {
        const int bytes = static_cast<int>(tc_smem(TC_MAX_SP, TC_MAX_SP).total);
        attr_err = cudaFuncSetAttribute(attention_fwd_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
        if (attr_err == cudaSuccess)
            attr_err = cudaFuncSetAttribute(attention_fwd_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    });
    UC2_REQUIRE(attr_err == cudaSuccess, UC2_ERR_CUDA, "attention_fwd_tc: cudaFuncSetAttribute failed: %s",
                cudaGetErrorString(attr_err));
    TcParams p;
    p.mask = attn_mask;
    p.ctx = static_cast<bf16*>(ctx);
    p.lse = lse;
    p.B = B; p.S = S; p.SP = SP; p.items = B * NH;
    p.drop = DropCfg{drop_key, drop_thresh, drop_scale};
    p.prof = nullptr;
    static const bool prof_on = [] { const char* e = getenv("UC2_ATTN_PROF"); return e && e[0] == '1'; }();
    static long long* prof_buf = nullptr;
    const int prof_words = 2 * 148 * (FWD_THREADS / 32) * PROF_SLOTS;
    if (prof_on) {
        if (!prof_buf) cudaMalloc(&prof_buf, prof_words * sizeof(long long));
        cudaMemsetAsync(prof_buf, 0, prof_words * sizeof(long long), (cudaStream_t)stream);
        p.prof = prof_buf;
    }
    // SP <= 160: two CTAs per SM (256 TMEM columns and < 113 KB of shared memory each); above: one, with all 512
    const bool pair = L.tm_cols == 256u;
    const int slots = (pair ? 2 : 1) * num_sms();
    const int grid = p.items < slots ? p.items : slots;
    ProfScope prof((cudaStream_t)stream, 1, 4.0 * B * NH * (double)S * S * HD);
    const size_t smem = pair ? L.total : exclusive_smem(L.total);
    const cudaError_t e = drop_thresh ? launch_pdl(attention_fwd_tc_kernel<true>, dim3(grid), dim3(FWD_THREADS), smem,
                                                   (cudaStream_t)stream, 1, tmap, tq0, tq1, p)
                                      : launch_pdl(attention_fwd_tc_kernel<false>, dim3(grid), dim3(FWD_THREADS), smem,
                                                   (cudaStream_t)stream, 1, tmap, tq0, tq1, p);
    UC2_REQUIRE(e == cudaSuccess, UC2_ERR_CUDA, "attention_fwd_tc launch failed: %s", cudaGetErrorString(e));
    if (prof_on) {        // debug only: synchronises; prints the phase cycle counters of CTA 0 and of the last CTA
        static int printed = 0;
        cudaStreamSynchronize((cudaStream_t)stream);
        if (printed++ < 4) {
            const int W = FWD_THREADS / 32;
            static long long host[2 * 148 * (FWD_THREADS / 32) * PROF_SLOTS];
            cudaMemcpy(host, prof_buf, prof_words * sizeof(long long), cudaMemcpyDeviceToHost);
            for (int c : {0, grid / 2, grid - 1})
                for (int w = 0; w < W; ++w) {
                    fprintf(stderr, "attn_fwd_tc prof B=%d S=%d drop=%u cta %d warp %d:", B, S, drop_thresh, c, w);
                    for (int k = 0; k < PROF_SLOTS; ++k) fprintf(stderr, " %lld", host[((long long)c * W + w) * PROF_SLOTS + k]);
                    fprintf(stderr, "\n");
                }
        }
    }
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
    UC2_REQUIRE(SP <= TC_BWD_MAX_SP, UC2_ERR_UNSUPPORTED, "attention_bwd_tc: S=%d > %d", S, TC_BWD_MAX_SP);
    UC2_REQUIRE(aligned16(qkv) && aligned16(ctx) && aligned16(dctx) && (reinterpret_cast<uintptr_t>(dqkv) & 31) == 0,
                UC2_ERR_ARG, "attention_bwd_tc: qkv / ctx / dctx must be 16-byte and dqkv 32-byte aligned");
    CUtensorMap tq, tdo, to;
    if (int rc = make_tmap(&tq, qkv, (long long)B * S, QKV_LD, QKV_LD, SP)) return rc;
    if (int rc = make_tmap(&tdo, dctx, (long long)B * S, HID, HID, SP)) return rc;
    if (int rc = make_tmap(&to, ctx, (long long)B * S, HID, HID, SP)) return rc;
    static const int blocks = [] { const char* e = getenv("UC2_ATTN_TC_BWD_BLOCKS"); return (e && e[0] == '2') ? 2 : 0; }();
    const TcBwdSmem L = tc_bwd_smem(S, SP, blocks);
    static std::once_flag once;
    static cudaError_t attr_err = cudaSuccess;
    static int nsplit = 4;
    std::call_once(once, [] {
        const char* e = getenv("UC2_ATTN_TC_BWD_SPLIT");          // tuning knob: 4 (default) or 2 warps per lane quarter
        if (e && e[0] == '2') nsplit = 2;
        const uint32_t b_wide = tc_bwd_smem(TC_BWD_MAX_SP, TC_BWD_MAX_SP).total, b_narrow = tc_bwd_smem(160, 160).total;
        const int bytes = static_cast<int>(b_wide > b_narrow ? b_wide : b_narrow);
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
    p.blocks = blocks;
    const int grid = p.items < num_sms() ? p.items : num_sms();
    p.prof = nullptr;
    static const bool prof_on = [] { const char* e = getenv("UC2_ATTN_PROF"); return e && e[0] == '1'; }();
    static long long* prof_buf = nullptr;
    const int prof_words = 148 * 32 * PROF_SLOTS;
    if (prof_on) {
        if (!prof_buf) cudaMalloc(&prof_buf, prof_words * sizeof(long long));
        cudaMemsetAsync(prof_buf, 0, prof_words * sizeof(long long), (cudaStream_t)stream);
        p.prof = prof_buf;
    }
    ProfScope prof((cudaStream_t)stream, 1, 10.0 * B * NH * (double)S * S * HD);
    const cudaError_t e =
        nsplit == 4 ? launch_pdl(attention_bwd_tc_kernel<4>, dim3(grid), dim3(bwd_threads(4)), exclusive_smem(L.total),
                                 (cudaStream_t)stream, 1, tq, tdo, to, p)
                    : launch_pdl(attention_bwd_tc_kernel<2>, dim3(grid), dim3(bwd_threads(2)), exclusive_smem(L.total),
                                 (cudaStream_t)stream, 1, tq, tdo, to, p);
    UC2_REQUIRE(e == cudaSuccess, UC2_ERR_CUDA, "attention_bwd_tc launch failed: %s", cudaGetErrorString(e));
    if (prof_on) {        // debug only: synchronises; prints the phase cycle counters of two CTAs
        static int printed = 0;
        cudaStreamSynchronize((cudaStream_t)stream);
        if (printed++ < 4) {
            const int W = bwd_threads(nsplit) / 32;
            static long long host[148 * 32 * PROF_SLOTS];
            cudaMemcpy(host, prof_buf, prof_words * sizeof(long long), cudaMemcpyDeviceToHost);
            for (int c : {0, grid - 1})
                for (int w = 0; w < W; ++w) {
                    fprintf(stderr, "attn_bwd_tc prof B=%d S=%d drop=%u cta %d warp %d:", B, S, drop_thresh, c, w);
                    for (int k = 0; k < PROF_SLOTS; ++k) fprintf(stderr, " %lld", host[((long long)c * W + w) * PROF_SLOTS + k]);
                    fprintf(stderr, "\n");
                }
        }
    }
    return check_last("attention_bwd_tc_kernel");
}
