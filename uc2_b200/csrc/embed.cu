// Text + region embeddings fused with the gather_index pack (forward and backward).
// Replaces UniterTextEmbeddings / VLXLMRTextEmbeddings (model/model.py:304-335, 987-1001),
// create_position_ids_from_input_ids (280-290), Uniter/VLXLMRImageEmbeddings (352-364, 1017-1028)
// after the img_linear GEMM, and the cat + gather of _compute_img_txt_embeddings (412-425).
// HBM-bound: one warp per OUTPUT row (b, j) of the packed [B, S, 768] sequence; the row pulls its
// source (text token or region) through gather_index, so pad columns that alias real rows
//  are reproduced exactly.
#include "common.cuh"

namespace uc2 {
namespace {

constexpr int VPL = 24;          // values per lane: 768 / 32
constexpr int EMB_WARPS = 8;

// element e = i*256 + lane*8 + k  (i < 3, k < 8): each lane owns 3 runs of 8 contiguous columns
__device__ __forceinline__ int col_of(int lane, int i) { return i * 256 + lane * 8; }

__device__ __forceinline__ void load_row_f32(const float* row, int lane, float* v) {
#pragma unroll
    for (int i = 0; i < 3; ++i) load8_f32(row + col_of(lane, i), v + 8 * i);
}
__device__ __forceinline__ void load_row_bf16(const bf16* row, int lane, float* v) {
#pragma unroll
    for (int i = 0; i < 3; ++i) load8_bf16(row + col_of(lane, i), v + 8 * i);
}
__device__ __forceinline__ void store_row_bf16(bf16* row, int lane, const float* v) {
#pragma unroll
    for (int i = 0; i < 3; ++i) store8_bf16(row + col_of(lane, i), v + 8 * i);
}

// two-pass LayerNorm statistics over the warp's 768 values
__device__ __forceinline__ void ln_stats(const float* v, float eps, float& mean, float& rstd) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) s += v[i];
    mean = warp_sum(s) * (1.0f / HID);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) { const float d = v[i] - mean; q += d * d; }
    rstd = rsqrtf(warp_sum(q) * (1.0f / HID) + eps);
}
__device__ __forceinline__ void ln_apply(float* v, const float* gamma, const float* beta, int lane, float eps) {
    float mean, rstd;
    ln_stats(v, eps, mean, rstd);
    float g[VPL], b[VPL];
    load_row_f32(gamma, lane, g);
    load_row_f32(beta, lane, b);
#pragma unroll
    for (int i = 0; i < VPL; ++i) v[i] = (v[i] - mean) * rstd * g[i] + b[i];
}

struct EmbedParams {
    int B, T, R, S;                 // S = rows per sample of the output
    int mode;                       // 0 joint (gather_index), 1 text only (S == T), 2 image only (S == R)
    const long long* input_ids;     // [B,T]
    const long long* position_ids;  // [pos_rows,T] or null (derive: VLXLMR)
    int pos_rows;                   // 1 (broadcast) or B
    const long long* gather_index;  // [B,S]
    int word_pad_id;                // token id that is "padding" for derived positions / grad skipping
    int pos_pad_id;                 // position_embeddings padding_idx or -1
    const float* word_emb; const float* pos_emb; const float* type_emb;   // fp32 master tables
    const float* ln_w; const float* ln_b;                                 // embeddings.LayerNorm
    const float* y_img;             // [B*R,768] fp32 = img_linear(feat (+mask emb)) incl. bias
    const float* pos_feat;          // [B*R,7]
    const float* img_ln_w; const float* img_ln_b;
    const float* pos_w; const float* pos_b;          // pos_linear [768,7], [768]
    const float* pos_ln_w; const float* pos_ln_b;
    const float* fin_ln_w; const float* fin_ln_b;    // img_embeddings.LayerNorm
    float eps;
    int vocab, max_pos;
    DropCfg drop;                   // dropout on the embedding outputs (model.py:334, 363), indexed by SOURCE row
};

struct Src { int is_img; int idx; };   // idx: token column t or region r

// Embedding dropout acts on the text / image embedding tensors BEFORE the gather_index pack, so the mask is a
// function of the source row b * (T + R) + (t | T + r): packed pad columns that alias a real row share its mask.
__device__ __forceinline__ void embed_dropout(const EmbedParams& p, int b, const Src& s, int lane, float* v) {
    if (p.drop.thresh == 0) return;
    const uint32_t src_row = (uint32_t)b * (uint32_t)(p.T + p.R) + (uint32_t)(s.is_img ? p.T + s.idx : s.idx);
#pragma unroll
    for (int i = 0; i < VPL; i += 2) {
        const uint32_t idx = src_row * HID + col_of(lane, i >> 3) + (i & 7);
        bool k0, k1;
        drop_keep2(p.drop.key, idx, p.drop.thresh, k0, k1);
        v[i] = k0 ? v[i] * p.drop.scale : 0.f;
        v[i + 1] = k1 ? v[i + 1] * p.drop.scale : 0.f;
    }
}

__device__ __forceinline__ Src resolve(const EmbedParams& p, int b, int j) {
    Src s;
    if (p.mode == 1) { s.is_img = 0; s.idx = j; return s; }
    if (p.mode == 2) { s.is_img = 1; s.idx = j; return s; }
    const long long g = p.gather_index[(long long)b * p.S + j];
    // an index outside the concatenation [text; regions] would read (and, in the backward, write) outside the batch:
    // fail the launch like nn.Embedding / torch.gather's device assert does
    if (g < 0 || g >= p.T + p.R) asm volatile("trap;");
    if (g < p.T) { s.is_img = 0; s.idx = (int)g; } else { s.is_img = 1; s.idx = (int)(g - p.T); }
    return s;
}

// position id of text column t: given, or cumsum over non-pad tokens (model.py:280-290)
__device__ __forceinline__ int text_position(const EmbedParams& p, int b, int t, int lane, long long id) {
    if (p.position_ids) return (int)p.position_ids[(long long)(p.pos_rows == 1 ? 0 : b) * p.T + t];
    if (id == p.word_pad_id) return p.word_pad_id;
    int cnt = 0;
    const long long* row = p.input_ids + (long long)b * p.T;
    for (int base = 0; base <= t; base += 32) {
        const int c = base + lane;
        const bool np = (c <= t) && (row[c] != p.word_pad_id);
        cnt += __popc(__ballot_sync(0xffffffffu, np));
    }
    return cnt + p.word_pad_id;
}

__device__ __forceinline__ void text_presum(const EmbedParams& p, long long id, int pos, int lane, float* v) {
    // token / position ids index the parameter (forward) and gradient (backward) tables directly: out of range = trap
    if (id < 0 || id >= p.vocab || pos < 0 || pos >= p.max_pos) asm volatile("trap;");
    float a[VPL];
    load_row_f32(p.word_emb + id * HID, lane, v);
    load_row_f32(p.pos_emb + (long long)pos * HID, lane, a);
#pragma unroll
    for (int i = 0; i < VPL; ++i) v[i] += a[i];
    load_row_f32(p.type_emb, lane, a);               // token type 0
#pragma unroll
    for (int i = 0; i < VPL; ++i) v[i] += a[i];
}

// pos_linear (7 -> 768) for the warp's columns
__device__ __forceinline__ void pos_linear(const EmbedParams& p, const float* f7, int lane, float* v) {
#pragma unroll
    for (int i = 0; i < 3; ++i) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int c = col_of(lane, i) + k;
            const float* w = p.pos_w + c * 7;
            float acc = p.pos_b[c];
#pragma unroll
            for (int d = 0; d < 7; ++d) acc += w[d] * f7[d];
            v[8 * i + k] = acc;
        }
    }
}

__global__ void __launch_bounds__(EMB_WARPS * 32)
embed_pack_fwd_kernel(const EmbedParams p, bf16* __restrict__ out, float* __restrict__ out32) {
    const int lane = threadIdx.x & 31;
    const long long row = (long long)blockIdx.x * EMB_WARPS + (threadIdx.x >> 5);
    if (row >= (long long)p.B * p.S) return;
    const int b = (int)(row / p.S), j = (int)(row % p.S);
    const Src s = resolve(p, b, j);
    float v[VPL];
    if (!s.is_img) {
        const long long id = p.input_ids[(long long)b * p.T + s.idx];
        const int pos = text_position(p, b, s.idx, lane, id);
        text_presum(p, id, pos, lane, v);
        ln_apply(v, p.ln_w, p.ln_b, lane, p.eps);
    } else {
        const long long ir = (long long)b * p.R + s.idx;
        load_row_f32(p.y_img + ir * HID, lane, v);
        ln_apply(v, p.img_ln_w, p.img_ln_b, lane, p.eps);
        float f7[7];
#pragma unroll
        for (int d = 0; d < 7; ++d) f7[d] = p.pos_feat[ir * 7 + d];
        float q[VPL];
        pos_linear(p, f7, lane, q);
        ln_apply(q, p.pos_ln_w, p.pos_ln_b, lane, p.eps);
        float t1[VPL];
        load_row_f32(p.type_emb + HID, lane, t1);    // token type 1 for every region (model.py:403-406)
#pragma unroll
        for (int i = 0; i < VPL; ++i) v[i] += q[i] + t1[i];
        ln_apply(v, p.fin_ln_w, p.fin_ln_b, lane, p.eps);
    }
    embed_dropout(p, b, s, lane, v);
    store_row_bf16(out + row * HID, lane, v);
    if (out32) {
#pragma unroll
        for (int i = 0; i < 3; ++i) store8_f32(out32 + row * HID + col_of(lane, i), v + 8 * i);
    }
}

// ------------------------------------------------------------------------------------------ backward
struct EmbedGrads {
    float* word_emb; float* pos_emb; float* type_emb;
    float* ln_w; float* ln_b;
    float* img_ln_w; float* img_ln_b;
    float* pos_w; float* pos_b;
    float* pos_ln_w; float* pos_ln_b;
    float* fin_ln_w; float* fin_ln_b;
    float* dy_img;      // [B*R,768] fp32, zero-initialised by the caller; accumulates d(img_linear out)
};

// Column accumulators of the backward kernel live in shared memory, one PRIVATE copy per warp, element i (0..23)
// of lane l at slot i * 32 + l: every lane only ever touches its own slots, so the updates are plain
// load-add-store (shared-memory atomics cost ~2 cycles per LANE; this kernel used to spend its time in them) and
// consecutive lanes hit consecutive banks.  Slot s holds column ((s >> 5) >> 3) * 256 + (s & 31) * 8 + ((s >> 5) & 7).
__device__ __forceinline__ int slot_col(int s) { return ((s >> 5) >> 3) * 256 + (s & 31) * 8 + ((s >> 5) & 7); }

// LayerNorm backward for one row held by the warp.  in: x (pre-LN), dy.  out: dx (into dy);
// accumulates dgamma/dbeta into this warp's accumulators.
__device__ __forceinline__ void ln_bwd_row(const float* x, float* dy, const float* gamma, int lane, float eps,
                                           float* acc_dgamma, float* acc_dbeta) {
    float mean, rstd;
    ln_stats(x, eps, mean, rstd);
    float g[VPL];
    load_row_f32(gamma, lane, g);
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
        const float xh = (x[i] - mean) * rstd;
        acc_dgamma[i * 32 + lane] += dy[i] * xh;
        acc_dbeta[i * 32 + lane] += dy[i];
        const float gd = g[i] * dy[i];
        s1 += gd;
        s2 += gd * xh;
    }
    s1 = warp_sum(s1) * (1.0f / HID);
    s2 = warp_sum(s2) * (1.0f / HID);
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
        const float xh = (x[i] - mean) * rstd;
        dy[i] = rstd * (g[i] * dy[i] - s1 - xh * s2);
    }
}

__device__ __forceinline__ void red_row(float* dst, int lane, const float* v) {
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        float* d = dst + col_of(lane, i);
        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(d), "f"(v[8 * i]), "f"(v[8 * i + 1]),
                     "f"(v[8 * i + 2]), "f"(v[8 * i + 3]) : "memory");
        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(d + 4), "f"(v[8 * i + 4]),
                     "f"(v[8 * i + 5]), "f"(v[8 * i + 6]), "f"(v[8 * i + 7]) : "memory");
    }
}

// shared accumulators (each 768 floats):
//  0 ln_w 1 ln_b 2 type0 | 3 fin_w 4 fin_b 5 type1 6 img_w 7 img_b 8 posln_w 9 posln_b 10 pos_b 11..17 pos_w[:,d]
constexpr int N_ACC = 18;

constexpr int EMB_BWD_WARPS = 4;      // 4 private accumulator sets x 18 x 768 fp32 = 216 KB of shared memory

__global__ void __launch_bounds__(EMB_BWD_WARPS * 32)
embed_pack_bwd_kernel(const EmbedParams p, const bf16* __restrict__ dout, const EmbedGrads g) {
    extern __shared__ float acc_all[];
    for (int i = threadIdx.x; i < EMB_BWD_WARPS * N_ACC * HID; i += blockDim.x) acc_all[i] = 0.f;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    float* acc = acc_all + (threadIdx.x >> 5) * N_ACC * HID;     // this warp's private accumulators
    const long long nrows = (long long)p.B * p.S;
    for (long long row = (long long)blockIdx.x * EMB_BWD_WARPS + (threadIdx.x >> 5); row < nrows;
         row += (long long)gridDim.x * EMB_BWD_WARPS) {
        const int b = (int)(row / p.S), j = (int)(row % p.S);
        const Src s = resolve(p, b, j);
        float dy[VPL];
        load_row_bf16(dout + row * HID, lane, dy);
        float nz = 0.f;
#pragma unroll
        for (int i = 0; i < VPL; ++i) nz += fabsf(dy[i]);
        if (warp_sum(nz) == 0.f) continue;       // rows that received no gradient (pads) add nothing
        embed_dropout(p, b, s, lane, dy);
        float x[VPL];
        if (!s.is_img) {
            const long long id = p.input_ids[(long long)b * p.T + s.idx];
            const int pos = text_position(p, b, s.idx, lane, id);
            text_presum(p, id, pos, lane, x);
            ln_bwd_row(x, dy, p.ln_w, lane, p.eps, acc + 0 * HID, acc + 1 * HID);
#pragma unroll
            for (int i = 0; i < VPL; ++i) acc[2 * HID + i * 32 + lane] += dy[i];
            // nn.Embedding(padding_idx) never receives gradient on its padding row
            if (id != p.word_pad_id) red_row(g.word_emb + id * HID, lane, dy);
            if (pos != p.pos_pad_id) red_row(g.pos_emb + (long long)pos * HID, lane, dy);
        } else {
            const long long ir = (long long)b * p.R + s.idx;
            // recompute the three pre-LN quantities
            float yi[VPL], a[VPL], q[VPL], qn[VPL], f7[7];
            load_row_f32(p.y_img + ir * HID, lane, yi);
#pragma unroll
            for (int i = 0; i < VPL; ++i) a[i] = yi[i];
            ln_apply(a, p.img_ln_w, p.img_ln_b, lane, p.eps);
#pragma unroll
            for (int d = 0; d < 7; ++d) f7[d] = p.pos_feat[ir * 7 + d];
            pos_linear(p, f7, lane, q);
#pragma unroll
            for (int i = 0; i < VPL; ++i) qn[i] = q[i];
            ln_apply(qn, p.pos_ln_w, p.pos_ln_b, lane, p.eps);
            load_row_f32(p.type_emb + HID, lane, x);
#pragma unroll
            for (int i = 0; i < VPL; ++i) x[i] += a[i] + qn[i];
            // final LN
            ln_bwd_row(x, dy, p.fin_ln_w, lane, p.eps, acc + 3 * HID, acc + 4 * HID);
#pragma unroll
            for (int i = 0; i < VPL; ++i) acc[5 * HID + i * 32 + lane] += dy[i];
            // branch: img LN
            float d1[VPL];
#pragma unroll
            for (int i = 0; i < VPL; ++i) d1[i] = dy[i];
            ln_bwd_row(yi, d1, p.img_ln_w, lane, p.eps, acc + 6 * HID, acc + 7 * HID);
            red_row(g.dy_img + ir * HID, lane, d1);
            // branch: pos LN -> pos_linear
            ln_bwd_row(q, dy, p.pos_ln_w, lane, p.eps, acc + 8 * HID, acc + 9 * HID);
#pragma unroll
            for (int i = 0; i < VPL; ++i) {
                acc[10 * HID + i * 32 + lane] += dy[i];
#pragma unroll
                for (int d = 0; d < 7; ++d) acc[(11 + d) * HID + i * 32 + lane] += dy[i] * f7[d];
            }
        }
    }
    __syncthreads();
    for (int sl = threadIdx.x; sl < HID; sl += blockDim.x) {
        const int c = slot_col(sl);
        auto flush = [&](int k, float* dst) {
            float v = 0.f;
#pragma unroll
            for (int w = 0; w < EMB_BWD_WARPS; ++w) v += acc_all[(w * N_ACC + k) * HID + sl];
            if (v != 0.f) atomicAdd(dst, v);
        };
        if (p.mode != 2) {
            flush(0, g.ln_w + c); flush(1, g.ln_b + c); flush(2, g.type_emb + c);
        }
        if (p.mode != 1) {
            flush(3, g.fin_ln_w + c); flush(4, g.fin_ln_b + c); flush(5, g.type_emb + HID + c);
            flush(6, g.img_ln_w + c); flush(7, g.img_ln_b + c);
            flush(8, g.pos_ln_w + c); flush(9, g.pos_ln_b + c); flush(10, g.pos_b + c);
#pragma unroll
            for (int d = 0; d < 7; ++d) flush(11 + d, g.pos_w + c * 7 + d);
        }
    }
}

// img_feat fp32 (+ mask_embedding row 1 on masked regions, model.py:352-356) -> bf16 GEMM operand
__global__ void img_prep_kernel(const float* __restrict__ feat, const unsigned char* __restrict__ masks,
                                const float* __restrict__ mask_row1, bf16* __restrict__ out, long long rows,
                                int dim) {
    const long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 8;
    if (i >= rows * dim) return;
    const long long r = i / dim;
    const int c = (int)(i % dim);
    float v[8];
    load8_f32(feat + i, v);
    if (masks && masks[r]) {
        float m[8];
        load8_f32(mask_row1 + c, m);
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] += m[k];
    }
    store8_bf16(out + i, v);
}

// dy_img fp32 [rows,768] -> bf16 copy for the wgrad GEMM, + column sums (img_linear.bias grad)
// + sum over masked rows (feeds the mask_embedding gradient)
__global__ void __launch_bounds__(256)
img_grad_finish_kernel(const float* __restrict__ dy, const unsigned char* __restrict__ masks, bf16* __restrict__ out,
                       float* __restrict__ dbias, float* __restrict__ masked_sum, long long rows) {
    __shared__ float acc[2 * HID];
    for (int i = threadIdx.x; i < 2 * HID; i += blockDim.x) acc[i] = 0.f;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    for (long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5); row < rows; row += (long long)gridDim.x * 8) {
        float v[VPL];
        load_row_f32(dy + row * HID, lane, v);
        store_row_bf16(out + row * HID, lane, v);
        const bool mk = masks && masks[row];
#pragma unroll
        for (int i = 0; i < VPL; ++i) {
            const int c = col_of(lane, i >> 3) + (i & 7);
            if (v[i] != 0.f) {
                atomicAdd(acc + c, v[i]);
                if (mk) atomicAdd(acc + HID + c, v[i]);
            }
        }
    }
    __syncthreads();
    for (int c = threadIdx.x; c < HID; c += blockDim.x) {
        if (acc[c] != 0.f) atomicAdd(dbias + c, acc[c]);
        if (masked_sum && acc[HID + c] != 0.f) atomicAdd(masked_sum + c, acc[HID + c]);
    }
}

// out[n] += sum_k v[k] * W[k][n]   (W fp32 [K,N]); used for d mask_embedding[1] = masked_sum @ img_linear.weight
__global__ void vecmat_acc_kernel(const float* __restrict__ v, const float* __restrict__ W, float* __restrict__ out,
                                  int K, int N) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    float acc = 0.f;
    for (int k = 0; k < K; ++k) acc += v[k] * W[(long long)k * N + n];
    atomicAdd(out + n, acc);
}

int fill_params(EmbedParams& p, const uc2_embed_args& a) {
    UC2_REQUIRE(a.B > 0 && a.S > 0, UC2_ERR_ARG, "embed: bad shape");
    UC2_REQUIRE(a.hidden == HID, UC2_ERR_UNSUPPORTED, "embed: kernels are specialised for hidden=768");
    p.B = a.B; p.T = a.T; p.R = a.R; p.S = a.S; p.mode = a.mode;
    p.input_ids = a.input_ids; p.position_ids = a.position_ids; p.pos_rows = a.position_rows;
    p.gather_index = a.gather_index; p.word_pad_id = a.word_pad_id; p.pos_pad_id = a.pos_pad_id;
    p.word_emb = a.word_emb; p.pos_emb = a.pos_emb; p.type_emb = a.type_emb; p.ln_w = a.ln_w; p.ln_b = a.ln_b;
    p.y_img = a.y_img; p.pos_feat = a.img_pos_feat; p.img_ln_w = a.img_ln_w; p.img_ln_b = a.img_ln_b;
    p.pos_w = a.pos_w; p.pos_b = a.pos_b; p.pos_ln_w = a.pos_ln_w; p.pos_ln_b = a.pos_ln_b;
    p.fin_ln_w = a.fin_ln_w; p.fin_ln_b = a.fin_ln_b; p.eps = a.eps; p.vocab = a.vocab; p.max_pos = a.max_pos;
    p.drop.key = a.drop_key; p.drop.thresh = a.drop_thresh; p.drop.scale = a.drop_scale;
    UC2_REQUIRE(a.drop_thresh < 65536u, UC2_ERR_ARG, "embed_pack: drop_thresh must be < 65536");
    if (a.mode == 0) UC2_REQUIRE(a.gather_index, UC2_ERR_ARG, "embed: joint mode needs gather_index");
    if (a.mode != 2) UC2_REQUIRE(a.input_ids && a.word_emb && a.pos_emb && a.type_emb && a.ln_w && a.ln_b,
                                 UC2_ERR_ARG, "embed: text inputs missing");
    if (a.mode != 1) UC2_REQUIRE(a.y_img && a.img_pos_feat && a.img_ln_w && a.pos_w && a.pos_ln_w && a.fin_ln_w &&
                                 a.type_emb, UC2_ERR_ARG, "embed: image inputs missing");
    if (a.mode == 1) UC2_REQUIRE(a.S == a.T, UC2_ERR_ARG, "embed: text-only mode needs S == T");
    if (a.mode == 2) UC2_REQUIRE(a.S == a.R, UC2_ERR_ARG, "embed: image-only mode needs S == R");
    return UC2_OK;
}

}  // namespace
}  // namespace uc2

using namespace uc2;

extern "C" UC2_API int uc2_img_prep(const float* img_feat, const unsigned char* img_masks, const float* mask_row1,
                                    void* out_bf16, long long rows, int dim, void* stream) {
    if (int rc = require_sm100()) return rc;
    UC2_REQUIRE(img_feat && out_bf16 && rows > 0 && dim % 8 == 0, UC2_ERR_ARG, "img_prep: bad args");
    UC2_REQUIRE(!img_masks || mask_row1, UC2_ERR_ARG, "img_prep: masks need the mask embedding row");
    const long long n8 = rows * dim / 8;
    img_prep_kernel<<<(unsigned)((n8 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        img_feat, img_masks, mask_row1, (bf16*)out_bf16, rows, dim);
    return check_last("img_prep_kernel");
}

extern "C" UC2_API int uc2_embed_pack_fwd(const uc2_embed_args* a, void* out_bf16, float* out_f32, void* stream) {
    if (int rc = require_sm100()) return rc;
    UC2_REQUIRE(a && out_bf16, UC2_ERR_ARG, "embed_pack_fwd: null");
    EmbedParams p;
    if (int rc = fill_params(p, *a)) return rc;
    const long long rows = (long long)p.B * p.S;
    embed_pack_fwd_kernel<<<(unsigned)((rows + EMB_WARPS - 1) / EMB_WARPS), EMB_WARPS * 32, 0,
                            (cudaStream_t)stream>>>(p, (bf16*)out_bf16, out_f32);
    return check_last("embed_pack_fwd_kernel");
}

extern "C" UC2_API int uc2_embed_pack_bwd(const uc2_embed_args* a, const void* dout_bf16,
                                          const uc2_embed_grads* gr, void* stream) {
    if (int rc = require_sm100()) return rc;
    UC2_REQUIRE(a && dout_bf16 && gr, UC2_ERR_ARG, "embed_pack_bwd: null");
    EmbedParams p;
    if (int rc = fill_params(p, *a)) return rc;
    EmbedGrads g;
    g.word_emb = gr->word_emb; g.pos_emb = gr->pos_emb; g.type_emb = gr->type_emb; g.ln_w = gr->ln_w;
    g.ln_b = gr->ln_b; g.img_ln_w = gr->img_ln_w; g.img_ln_b = gr->img_ln_b; g.pos_w = gr->pos_w;
    g.pos_b = gr->pos_b; g.pos_ln_w = gr->pos_ln_w; g.pos_ln_b = gr->pos_ln_b; g.fin_ln_w = gr->fin_ln_w;
    g.fin_ln_b = gr->fin_ln_b; g.dy_img = gr->dy_img;
    if (p.mode != 2) UC2_REQUIRE(g.word_emb && g.pos_emb && g.type_emb && g.ln_w && g.ln_b, UC2_ERR_ARG,
                                 "embed_pack_bwd: text grads missing");
    if (p.mode != 1) UC2_REQUIRE(g.img_ln_w && g.img_ln_b && g.pos_w && g.pos_b && g.pos_ln_w && g.pos_ln_b &&
                                 g.fin_ln_w && g.fin_ln_b && g.dy_img && g.type_emb, UC2_ERR_ARG,
                                 "embed_pack_bwd: image grads missing");
    static bool attr_set = false;
    const int smem = EMB_BWD_WARPS * N_ACC * HID * (int)sizeof(float);
    if (!attr_set) {
        UC2_CUDA(cudaFuncSetAttribute(embed_pack_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        attr_set = true;
    }
    const long long rows = (long long)p.B * p.S;
    long long blocks = (rows + EMB_BWD_WARPS - 1) / EMB_BWD_WARPS;
    const long long cap = num_sms();              // the accumulators take most of an SM's shared memory
    if (blocks > cap) blocks = cap;
    embed_pack_bwd_kernel<<<(unsigned)blocks, EMB_BWD_WARPS * 32, smem, (cudaStream_t)stream>>>(
        p, (const bf16*)dout_bf16, g);
    return check_last("embed_pack_bwd_kernel");
}

extern "C" UC2_API int uc2_img_grad_finish(const float* dy_img, const unsigned char* img_masks, void* dy_bf16,
                                           float* dbias, float* masked_sum, long long rows, void* stream) {
    if (int rc = require_sm100()) return rc;
    UC2_REQUIRE(dy_img && dy_bf16 && dbias && rows > 0, UC2_ERR_ARG, "img_grad_finish: bad args");
    long long blocks = (rows + 7) / 8;
    if (blocks > 2LL * num_sms()) blocks = 2LL * num_sms();
    img_grad_finish_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(dy_img, img_masks, (bf16*)dy_bf16,
                                                                               dbias, masked_sum, rows);
    return check_last("img_grad_finish_kernel");
}

extern "C" UC2_API int uc2_vecmat_acc(const float* v, const float* W, float* out, int K, int N, void* stream) {
    if (int rc = require_sm100()) return rc;
    UC2_REQUIRE(v && W && out && K > 0 && N > 0, UC2_ERR_ARG, "vecmat_acc: bad args");
    vecmat_acc_kernel<<<(N + 127) / 128, 128, 0, (cudaStream_t)stream>>>(v, W, out, K, N);
    return check_last("vecmat_acc_kernel");
}
