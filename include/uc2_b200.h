/* uc2_b200 -- C ABI of the B200 (sm_100a) kernels behind the UC2 cross-modal encoder path.
 *
 * The reference (zmykevin/UC2) is pure Python/PyTorch and has no FFI of its own; each entry
 * point below names the reference function (file:line under the reference tree) whose device
 * work it replaces.  Conventions:
 *   - plain pointers + sizes, no torch types; every pointer is a DEVICE pointer unless noted;
 *   - the caller owns every buffer (inputs, outputs, saved-for-backward, workspace);
 *   - every function returns 0 on success or a negative UC2_ERR_* code and never throws;
 *     uc2_last_error() returns a message for the calling thread's last failure;
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it, nothing syncs;
 *   - bf16 tensors are row-major uint16 storage (__nv_bfloat16), fp32 tensors are float.
 */
#ifndef UC2_B200_H
#define UC2_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(UC2_BUILD)
#define UC2_API __attribute__((visibility("default")))
#else
#define UC2_API
#endif

#define UC2_OK 0
#define UC2_ERR_ARG (-1)      /* bad shape / null / misaligned pointer */
#define UC2_ERR_ARCH (-2)     /* device is not sm_100 */
#define UC2_ERR_CUDA (-3)     /* a CUDA runtime/driver call failed */
#define UC2_ERR_UNSUPPORTED (-4)

UC2_API const char* uc2_last_error(void);
UC2_API int uc2_version(void);
/* Number of kernels this library has launched in the calling process (bench.py gpu_launches). */
UC2_API long long uc2_launch_count(void);

/* Optional per-launch timing with CUDA events on the launching stream (used by bench.py for the roofline
 * numbers; off by default).  collect() synchronises, sums duration [ms], algorithmic work [FLOP] and launch
 * count per kernel kind (0 = GEMM, 1 = attention, 2 = other) into arrays of `kinds` entries and clears. */
UC2_API int uc2_profile_enable(int on);
UC2_API int uc2_profile_collect(double* ms_by_kind, double* work_by_kind, int* launches_by_kind, int kinds);

/* ---------------------------------------------------------------------------------------------
 * Dense layers: tcgen05/TMEM GEMM fed by TMA.
 * Replaces every nn.Linear on the path: model/layer.py:76-78 (Q,K,V), 112 (attention out),
 * 140 (FFN1), 153 (FFN2); model/model.py:359 (img_linear), heads 1153-1169; and their autograd
 * backward (dgrad / wgrad).
 *
 *   D[M,N] = A[M,K] * B[N,K]^T   (+ epilogue)
 * a_mn != 0 : A is stored transposed, i.e. a[K][M] row-major with leading dimension lda (wgrad).
 * b_mn != 0 : B is stored transposed, i.e. b[K][N] row-major with leading dimension ldb (dgrad).
 * Epilogue (applied in this order, every pointer optional unless noted):
 *   acc += bias[n] (fp32)                       -- nn.Linear bias
 *   if out_pre : out_pre[m,n] = bf16(acc)       -- pre-activation copy (saved for backward)
 *   if act == UC2_ACT_GELU : acc = gelu_erf(acc)          (model/layer.py:31-37)
 *   if act == UC2_ACT_DGELU: acc *= gelu_erf'(aux[m,n])   (backward of the above; aux = pre-act)
 *   if act == UC2_ACT_TANH : acc = tanh(acc)              (model/layer.py:184)
 *   if drop_thresh : acc = keep(m,n) ? acc * drop_scale : 0  -- nn.Dropout before the residual add
 *   acc += residual[m,n] (bf16, or fp32 when residual_f32) -- BertSelfOutput / BertOutput residual
 *   out_bf16[m,n] = bf16(acc)  and/or  out_f32[m,n] (= or +=, see accumulate) acc
 * split_k > 1 splits K over CTAs and atomically accumulates into out_f32 (requires accumulate=1,
 * out_f32 only, no bias/act/residual): the wgrad path, which also gives gradient accumulation.
 */
#define UC2_ACT_NONE 0
#define UC2_ACT_GELU 1
#define UC2_ACT_DGELU 2
#define UC2_ACT_TANH 3

typedef struct {
    const void* a; long long lda; int a_mn;
    const void* b; long long ldb; int b_mn;
    int M, N, K;
    const float* bias;
    const void* residual; long long ld_res;
    const void* aux; long long ld_aux;
    int act;
    void* out_bf16; long long ld_out;
    void* out_pre; long long ld_pre;
    float* out_f32; long long ld_f32;
    int accumulate;   /* out_f32 += acc (atomic) instead of = */
    int split_k;      /* >= 1 */
    int block_n;      /* 0 = auto; else 64, 128 or 256 */
    int residual_f32; /* residual points at fp32 (the fp32 residual stream of the encoder) instead of bf16 */
    int ctas;         /* 0 = auto; 1 = one CTA per 128-row tile; 2 = CTA pair (cta_group::2) per 256-row tile */
    /* dropout on (acc + bias) after the activation and before the residual add: nn.Dropout of BertSelfOutput /
     * BertOutput (model/layer.py:113, 154).  keep(m, n) <=> (lowbias32((m * N + n) ^ drop_key) >> 16) >= drop_thresh;
     * kept values are multiplied by drop_scale.  drop_thresh = 0 disables it (uc2_b200/dropout.py has the rules). */
    unsigned int drop_key; unsigned int drop_thresh; float drop_scale;
    int tail_split;   /* 0 = auto (split the tiles of a partial last round into column slices); 1 = never */
    /* Cross-entropy statistics fused into the epilogue (the tied MLM decoder, model/layer.py:257-265 followed by
     * F.cross_entropy, model/model.py:592-596): with ce_stats != NULL the call must be  out_bf16 = A B^T + bias  and the
     * epilogue also writes, per row m and per 32-column chunk c, the pair (max, sum exp(z - max)) of the fp32 values
     * z = acc + bias of that chunk to ce_stats[(c * ld_ce + m) * 2 ..], and z[m, ce_labels[m]] to ce_tgt[m].
     * uc2_ce_stats_reduce turns them into lse / loss; no fp32 [M, N] tensor is ever materialised. */
    float* ce_stats; long long ld_ce; const long long* ce_labels; float* ce_tgt;
} uc2_gemm_args;

UC2_API int uc2_gemm_bf16(const uc2_gemm_args* args, void* stream);

/* Second half of the fused cross entropy: per row, merge the n_chunks (max, sum-exp) pairs the GEMM epilogue left in
 * stats ([n_chunks][ld_ce] pairs) into lse[row], and loss[row] = lse - tgt[row] (0 where targets[row] == ignore_index).
 * part is a scratch of ceil(n_chunks / 256) * rows pairs. */
UC2_API int uc2_ce_stats_reduce(const float* stats, long long ld_ce, int n_chunks, long long rows, const float* tgt,
                                const long long* targets, long long ignore_index, float* part, float* loss, float* lse,
                                void* stream);
/* d(logits) in place over the bf16 logits: z <- dloss[row] * (exp(z - lse[row]) - [col == target]) (0 for ignored rows) */
UC2_API int uc2_ce_bwd_inplace_bf16(void* logits_bf16, long long ld, long long rows, int C, const long long* targets,
                                    long long ignore_index, const float* dloss, const float* lse, void* stream);

/* Persistent kernels (GEMM, attention, LayerNorm backward, ...) size their grids for every SM of the device minus n
 * (even, <= 64; UC2_RESERVE_SMS in the environment sets the initial value).  Data-parallel training sets aside the SMs
 * it lets NCCL use, so that neither side waits for an SM the other holds.  Returns the previous value. */
UC2_API int uc2_reserve_sms(int n);

/* Tile order of the persistent GEMM workers.  0 (default): fixed round-robin, no per-launch overhead.  1: tiles are drawn
 * from a device counter, so a worker whose SM was busy when the grid started (a NCCL kernel of the overlapped gradient
 * exchange holds it) finds the work done and leaves instead of running its whole fixed share late.  Data-parallel
 * training (world size > 1) switches it on; UC2_GEMM_SCHED=dynamic|static in the environment sets the initial value.
 * Same results either way (split-K accumulation order aside).  Returns the previous value. */
UC2_API int uc2_gemm_sched_dynamic(int on);
/* Test support: n_ctas CTAs that each hold a whole SM (200 KB of shared memory) for `cycles` clocks on `stream` -- a
 * stand-in for a communication kernel that keeps SMs from the persistent kernels. */
UC2_API int uc2_debug_occupy_sms(int n_ctas, long long cycles, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Embeddings fused with the gather_index pack.
 * Replaces {VLXLMR,Uniter}TextEmbeddings.forward (model/model.py:304-335, 987-1001),
 * create_position_ids_from_input_ids (280-290), {VLXLMR,Uniter}ImageEmbeddings.forward
 * (352-364, 1017-1028; the img_linear GEMM itself goes through uc2_gemm_bf16) and the
 * cat+gather of _compute_img_txt_embeddings (412-425), plus their backward.
 * All parameter tensors are the fp32 masters.  Output rows: out[b, j] for j < S.
 *   mode 0: joint, out[b,j] = cat(txt,img)[b, gather_index[b,j]]
 *   mode 1: text only (S == T);  mode 2: image only (S == R)
 * position_ids == NULL derives positions from input_ids (VLXLMR); position_rows is 1 (broadcast
 * [1,T]) or B.  word_pad_id is the token treated as padding for derived positions AND the
 * word-embedding row that never receives gradient; pos_pad_id likewise for the position table
 * (-1: none).
 */
typedef struct {
    int B, T, R, S, mode, hidden;
    const long long* input_ids;
    const long long* position_ids; int position_rows;
    const long long* gather_index;
    int word_pad_id, pos_pad_id;
    const float* word_emb; const float* pos_emb; const float* type_emb;
    const float* ln_w; const float* ln_b;
    const float* y_img;          /* [B*R, 768] fp32: img_linear(img_feat (+mask emb)) incl. bias */
    const float* img_pos_feat;   /* [B*R, 7] */
    const float* img_ln_w; const float* img_ln_b;
    const float* pos_w; const float* pos_b;
    const float* pos_ln_w; const float* pos_ln_b;
    const float* fin_ln_w; const float* fin_ln_b;
    float eps;
    int vocab, max_pos;
    /* nn.Dropout on the text / image embeddings (model/model.py:334, 363), applied before the pack: element
     * (b, source row, col) is kept iff (lowbias32(((b * (T + R) + src) * 768 + col) ^ drop_key) >> 16) >= drop_thresh,
     * src = t for text, T + r for regions; drop_thresh = 0 disables it. */
    unsigned int drop_key; unsigned int drop_thresh; float drop_scale;
} uc2_embed_args;

typedef struct {                 /* fp32 gradient accumulators (+=), same shapes as the parameters */
    float* word_emb; float* pos_emb; float* type_emb;
    float* ln_w; float* ln_b;
    float* img_ln_w; float* img_ln_b;
    float* pos_w; float* pos_b;
    float* pos_ln_w; float* pos_ln_b;
    float* fin_ln_w; float* fin_ln_b;
    float* dy_img;               /* [B*R, 768] fp32 scratch, zeroed by the caller: d(img_linear out) */
} uc2_embed_grads;

/* img_feat fp32 [rows, dim] (+ mask_embedding.weight[1] where img_masks[row] != 0) -> bf16 */
UC2_API int uc2_img_prep(const float* img_feat, const unsigned char* img_masks, const float* mask_row1,
                         void* out_bf16, long long rows, int dim, void* stream);
/* out_bf16 [B*S,768] (tensor-core operand) and, optionally, out_f32 (start of the fp32 residual stream) */
UC2_API int uc2_embed_pack_fwd(const uc2_embed_args* a, void* out_bf16, float* out_f32, void* stream);
UC2_API int uc2_embed_pack_bwd(const uc2_embed_args* a, const void* dout_bf16, const uc2_embed_grads* g,
                               void* stream);
/* dy_img fp32 -> bf16 copy (wgrad operand), dbias += column sums, masked_sum += sum of masked rows */
UC2_API int uc2_img_grad_finish(const float* dy_img, const unsigned char* img_masks, void* dy_bf16, float* dbias,
                                float* masked_sum, long long rows, void* stream);
/* out[n] += sum_k v[k] W[k,n]  (d mask_embedding.weight[1] = masked_sum @ img_linear.weight) */
UC2_API int uc2_vecmat_acc(const float* v, const float* W, float* out, int K, int N, void* stream);

/* ---------------------------------------------------------------------------------------------
 * LayerNorm over rows of 768 (apex FusedLayerNorm call sites model/layer.py:108,149 and the
 * head transforms), forward and backward.  x/y/dy/dx are bf16 [rows, 768]; gamma/beta fp32.
 * Backward accumulates dgamma/dbeta (+=) and, if dbias != NULL, the column sums of dx (the bias
 * gradient of the Linear that produced x).
 */
/* x is bf16, or fp32 when x_is_f32 (the encoder's fp32 residual stream); y_bf16 and/or y_f32 are written. */
UC2_API int uc2_layernorm_fwd(const void* x, int x_is_f32, const float* gamma, const float* beta, float eps,
                              void* y_bf16, float* y_f32, long long rows, void* stream);
UC2_API int uc2_layernorm_bwd(const void* x, int x_is_f32, const void* dy, const float* gamma, float eps, void* dx,
                              float* dgamma, float* dbeta, float* dbias, long long rows, void* stream);
/* Same, for x = dropout(dense) + residual (training): also writes dx_masked = keep ? dx * drop_scale : 0 (the gradient
 * the dense branch sees; element index row * 768 + col under drop_key) and takes dbias from the masked gradient. */
UC2_API int uc2_layernorm_bwd_dropout(const void* x, int x_is_f32, const void* dy, const float* gamma, float eps,
                                      void* dx, float* dgamma, float* dbeta, float* dbias, long long rows,
                                      void* dx_masked, unsigned int drop_key, unsigned int drop_thresh,
                                      float drop_scale, void* stream);
/* out[c] += sum_r x[r, c] for bf16 x [rows, cols] with leading dimension ld (bias gradients) */
UC2_API int uc2_colsum_bf16(const void* x, long long ld, long long rows, int cols, float* out, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Key-padding-masked multi-head self-attention, flash style (no [B,12,S,S] tensor).
 * Replaces BertSelfAttention.forward model/layer.py:80-100 after the QKV projection.
 * qkv: bf16 [B*S, 2304] = [q | k | v], heads of 64 inside each third.  attn_mask: int64 [B,S]
 * of {0,1}; the additive mask is (1-m)*-10000 exactly as model/model.py:433-436.
 * ctx: bf16 [B*S, 768] (heads merged, i.e. the permute+view of layer.py:98-100 is free).
 * lse: fp32 [B,12,S] log-sum-exp saved for backward.
 */
UC2_API int uc2_attention_fwd(const void* qkv, const long long* attn_mask, void* ctx, float* lse, int B, int S,
                              void* stream);
/* dqkv: bf16 [B*S, 2304] (fully overwritten). delta_ws: fp32 [B,12,S] scratch (rowsum(dO * O); only written for
 * S > 160, shorter sequences compute it inside the backward kernel). */
UC2_API int uc2_attention_bwd(const void* qkv, const long long* attn_mask, const void* ctx, const void* dctx,
                              const float* lse, float* delta_ws, void* dqkv, int B, int S, void* stream);

/* Training forms with attention-probability dropout (model/layer.py:94): P is dropped and rescaled after the
 * softmax normalisation (the normaliser keeps every key).  Forward and backward regenerate the mask from
 * (drop_key, batch * 12 + head, query, key); drop_thresh = round(p * 65536) < 65536, drop_scale = 1 / (1 - p),
 * drop_thresh = 0 switches it off.  S <= 256.
 *
 * Which kernels run: by default the tcgen05 / TMEM kernels below for every shape they take (forward S <= 256,
 * backward S <= 160), the mma.sync kernels of csrc/attention.cu otherwise.  The two families draw the dropout mask
 * from different counter streams (tcgen05: one lowbias32 per 16 x 16 block of the [S, S] mask, then
 * e = h * CA^(q & 15) * CB^(k & 15) mod 2^32, kept iff e >= drop_thresh << 16 -- uc2_b200/dropout.py
 * attn_keep_mask_np is the host mirror; mma.sync: keep_mask_np over q * S + k), so with dropout on the forward only
 * goes to tcgen05 when the backward of that shape does too. */
UC2_API int uc2_attention_fwd_dropout(const void* qkv, const long long* attn_mask, void* ctx, float* lse, int B, int S,
                                      unsigned int drop_key, unsigned int drop_thresh, float drop_scale, void* stream);
UC2_API int uc2_attention_bwd_dropout(const void* qkv, const long long* attn_mask, const void* ctx, const void* dctx,
                                      const float* lse, float* delta_ws, void* dqkv, int B, int S,
                                      unsigned int drop_key, unsigned int drop_thresh, float drop_scale, void* stream);
/* The tcgen05 / TMEM attention kernels themselves (csrc/attention_tc.cu), same arguments and outputs: S = Q K^T and
 * O = P V on the 5th-generation tensor cores with the accumulators (and P, as the TMEM-A operand of P V) in tensor
 * memory, Q / K / V tiles by TMA.  Forward: S <= 256, ctx 32-byte aligned.  Backward: S <= 160, dqkv 32-byte aligned
 * and fully overwritten for the B*S rows; delta = rowsum(dO * O) is computed inside from the dO and O tiles, so there
 * is no workspace argument.  UC2_ERR_UNSUPPORTED above those lengths.  A mask row that is a prefix of ones (every
 * UC2 collate) only costs its valid keys: masked keys have probability exactly 0 in fp32 next to any valid key.
 * uc2_attention_tc_enable(0) (or UC2_ATTN_TCGEN05=0 in the environment) routes the public entry points
 * (uc2_attention_fwd*, _bwd*, uc2_encoder_fwd*, _bwd*) back to the mma.sync kernels; it returns the previous setting.
 * UC2_ATTN_TC_BWD_SPLIT=2 runs the backward with 8 instead of 16 element-wise warps (a tuning knob; same results). */
UC2_API int uc2_attention_fwd_tc(const void* qkv, const long long* attn_mask, void* ctx, float* lse, int B, int S,
                                 unsigned int drop_key, unsigned int drop_thresh, float drop_scale, void* stream);
UC2_API int uc2_attention_bwd_tc(const void* qkv, const long long* attn_mask, const void* ctx, const void* dctx,
                                 const float* lse, void* dqkv, int B, int S, unsigned int drop_key,
                                 unsigned int drop_thresh, float drop_scale, void* stream);
UC2_API int uc2_attention_tc_enable(int on);

/* ---------------------------------------------------------------------------------------------
 * The BertLayer stack: BertLayer.forward model/layer.py:159-170 applied num_hidden_layers times
 * ({VLXLMR,Uniter}Encoder.forward model/model.py:373-383 / 1038-1048) and its backward.
 * The arrays w / acts / grads are HOST arrays of n_layers structs holding DEVICE pointers.
 * Weights are the bf16 shadows ([out,in] row-major, Q/K/V concatenated along `out`), biases and
 * LayerNorm parameters the fp32 masters.  Activations are bf16 except lse (fp32 [B,12,S]).
 * For inference (save_for_bwd == 0) every layer may point at the same activation buffers and
 * `u` may be NULL.
 */
typedef struct {
    const void* w_qkv; const float* b_qkv;      /* [2304,768], [2304] */
    const void* w_o; const float* b_o;          /* attention.output.dense */
    const float* ln1_w; const float* ln1_b;     /* attention.output.LayerNorm (eps 1e-12) */
    const void* w_ffn1; const float* b_ffn1;    /* intermediate.dense [3072,768] */
    const void* w_ffn2; const float* b_ffn2;    /* output.dense [768,3072] */
    const float* ln2_w; const float* ln2_b;     /* output.LayerNorm (eps 1e-12) */
} uc2_layer_weights;

typedef struct {                                /* fp32 accumulators (+=), shapes as above */
    float* w_qkv; float* b_qkv; float* w_o; float* b_o; float* ln1_w; float* ln1_b;
    float* w_ffn1; float* b_ffn1; float* w_ffn2; float* b_ffn2; float* ln2_w; float* ln2_b;
} uc2_layer_grads;

typedef struct {
    void* qkv;    /* [M,2304] */
    void* ctx;    /* [M,768]  attention output, heads merged */
    float* lse;   /* [B,12,S] */
    void* z1;     /* [M,768]  fp32: O-proj + bias + residual (LayerNorm input) */
    void* h1;     /* [M,768]  attention block output */
    void* u;      /* [M,3072] FFN1 pre-activation */
    void* g;      /* [M,3072] gelu(u) */
    void* z2;     /* [M,768]  fp32: FFN2 + bias + residual */
    void* out;    /* [M,768]  layer output */
    /* dropout site keys of this layer (attention probabilities, attention output, FFN output); used when the
     * uc2_dropout argument of uc2_encoder_fwd / _bwd enables the site.  Living here, they stay attached to the
     * layer when backward runs over sub-ranges of layers. */
    unsigned int key_attn, key_out1, key_out2;
} uc2_layer_acts;

typedef struct {
    unsigned int attn_thresh; float attn_scale;       /* attention_probs_dropout_prob (model/layer.py:94) */
    unsigned int hidden_thresh; float hidden_scale;   /* hidden_dropout_prob (model/layer.py:113, 154) */
} uc2_dropout;

/* x_in: bf16 [M,768] embedding output; x_in_f32: the same rows in fp32 (start of the residual stream).
 * workspace: uc2_encoder_fwd_workspace_bytes(B,S) bytes, 256-byte aligned (rotating fp32 residual buffers). */
UC2_API size_t uc2_encoder_fwd_workspace_bytes(int B, int S);
UC2_API int uc2_encoder_fwd(const void* x_in, const float* x_in_f32, const long long* attn_mask, int B, int S,
                            int n_layers, const uc2_layer_weights* w, const uc2_layer_acts* acts, int save_for_bwd,
                            void* workspace, size_t workspace_bytes, void* stream);
/* Training forms: `drop` (may be NULL = no dropout) switches the three per-layer dropout sites on. */
UC2_API int uc2_encoder_fwd_dropout(const void* x_in, const float* x_in_f32, const long long* attn_mask, int B, int S,
                                    int n_layers, const uc2_layer_weights* w, const uc2_layer_acts* acts,
                                    int save_for_bwd, const uc2_dropout* drop, void* workspace, size_t workspace_bytes,
                                    void* stream);
UC2_API size_t uc2_encoder_bwd_workspace_bytes(int B, int S);
/* dout: bf16 [M,768] gradient of the last layer's output; dx_in: bf16 [M,768] gradient of x_in (written). */
UC2_API int uc2_encoder_bwd(const void* x_in, const long long* attn_mask, int B, int S, int n_layers,
                            const uc2_layer_weights* w, const uc2_layer_acts* acts, const uc2_layer_grads* grads,
                            const void* dout, void* dx_in, void* workspace, size_t workspace_bytes, void* stream);
UC2_API int uc2_encoder_bwd_dropout(const void* x_in, const long long* attn_mask, int B, int S, int n_layers,
                                    const uc2_layer_weights* w, const uc2_layer_acts* acts, const uc2_layer_grads* grads,
                                    const void* dout, void* dx_in, const uc2_dropout* drop, void* workspace,
                                    size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Heads and losses (warp-level kernels).
 */
/* Narrow Linear (N <= 8): itm_output Linear(768,2) model/model.py:476,698; rank_output Linear(768,1)
 * model/itm.py:19,42.  x fp32 [M,K] (row pitch ldx), W fp32 [N,K]; out fp32 [M,N]. */
UC2_API int uc2_narrow_linear_fwd(const float* x, long long ldx, const float* W, const float* b, float* out, int M,
                                  int N, int K, void* stream);
/* dx (fp32, row pitch lddx, may be NULL) = dy W;  dW += dy^T x;  db += colsum(dy) */
UC2_API int uc2_narrow_linear_bwd(const float* x, long long ldx, const float* W, const float* dy, float* dx,
                                  long long lddx, float* dW, float* db, int M, int N, int K, void* stream);
/* BertPooler tanh backward (model/layer.py:184): dpre = dy * (1 - y^2), written as bf16 */
UC2_API int uc2_tanh_bwd(const float* y, const float* dy, void* dpre_bf16, long long n, void* stream);
/* Triplet ranking loss, model/itm.py:43-53: scores [groups*sample_size] -> loss [groups, sample_size-1] */
UC2_API int uc2_rank_loss_fwd(const float* scores, float* loss, int groups, int sample_size, float margin, void* stream);
UC2_API int uc2_rank_loss_bwd(const float* scores, const float* dloss, float* dscores, int groups, int sample_size,
                              float margin, void* stream);
/* Row-wise log-softmax losses over fp32 logits [rows, C] (pitch ld), reduction 'none':
 *   kind 0: F.cross_entropy with int64 targets (model/model.py:593-595, 732, 770-772; ignore_index or -1)
 *   kind 1: F.kl_div(log_softmax(x), soft_targets) elementwise [rows, C] (model/model.py:763-766)
 * loss / dlogits / lse_out are optional outputs; dlogits needs dloss (same shape as loss). */
UC2_API int uc2_softmax_loss(const float* logits, long long ld, long long rows, int C, int kind,
                             const long long* targets, long long ignore_index, const float* soft_targets, float* loss,
                             const float* dloss, float* dlogits, float* lse_out, void* stream);
/* Large-vocabulary form of kind 0 for the tied MLM decoder (model/layer.py:257-265 + model/model.py:592-596): forward
 * is one pass per row (online max / sum) and also returns the row log-sum-exp; backward reads lse instead of
 * recomputing it and writes d(logits) as the bf16 matrix (row pitch ld_out elements) the decoder's dgrad / wgrad
 * GEMMs consume. */
UC2_API int uc2_ce_loss_fwd(const float* logits, long long ld, long long rows, int C, const long long* targets,
                            long long ignore_index, float* loss, float* lse, void* stream);
UC2_API int uc2_ce_loss_bwd_bf16(const float* logits, long long ld, long long rows, int C, const long long* targets,
                                 long long ignore_index, const float* dloss, const float* lse, void* dlogits_bf16,
                                 long long ld_out, void* stream);
/* F.mse_loss(reduction='none') model/model.py:684-685 and its backward */
UC2_API int uc2_mse(const float* pred, const float* tgt, float* loss, const float* dloss, float* dpred, long long n,
                    void* stream);
/* Order-exact masked-row compaction, _compute_masked_hidden model/model.py:653-657, without a host sync.
 * mask: bytes [B*mask_L] row-major; index[i] = flat position of the i-th set byte; *count = number set. */
UC2_API int uc2_mask_scan(const unsigned char* mask, long long n, int* index, int* count, int capacity, void* stream);
/* out[i,:] = src[(index[i] / mask_L) * src_S + index[i] % mask_L, :] for i < min(*count, capacity); rows of 768 bf16 */
UC2_API int uc2_gather_rows(const void* src, const int* index, const int* count, int mask_L, int src_S, void* out,
                            int capacity, void* stream);
UC2_API int uc2_scatter_rows_add(const void* dout, const int* index, const int* count, int mask_L, int src_S,
                                 void* dsrc, int capacity, void* stream);
/* dz = dy * gelu_erf'(pre), bf16 elementwise (backward of the head transforms, model/layer.py:258-260) */
UC2_API int uc2_dgelu_bf16(const void* dy, const void* pre, void* out, long long n, void* stream);
UC2_API int uc2_f32_to_bf16_2d(const float* x, long long ldx, void* y, long long ldy, long long rows, int cols,
                               void* stream);

/* ---------------------------------------------------------------------------------------------
 * WRA optimal transport: optimal_transport_dist model/ot.py:66-82 (cost_matrix_cosine 8-18, ipot
 * 32-63, trace 21-29) fused with the scatter un-pack of forward_itm model/model.py:703-716.
 * seq: bf16 [B,S,768] packed encoder output; ot_scatter: int64 [B,S] context position of each packed
 * row; text rows are context positions [0,M), image rows [tl, tl+N) with tl = input_ids.size(1);
 * txt_pad [B,M] / img_pad [B,N]: bytes, 1 = padding.  dist: fp32 [B].  C_save [B,N,M] / T_save
 * [B,N,M] fp32 are saved for backward.  Backward writes d(dist)/d(seq) (T detached, as ot.py:79-81)
 * into dseq (bf16 [B,S,768], zero-initialised by the caller).  Limits: M, N <= 128.
 */
UC2_API int uc2_ot_ipot_fwd(const void* seq, const long long* ot_scatter, const unsigned char* txt_pad,
                            const unsigned char* img_pad, int B, int S, int M, int N, int tl, float beta,
                            int iterations, int k, float* dist, float* C_save, float* T_save, void* stream);
UC2_API int uc2_ot_ipot_bwd(const void* seq, const long long* ot_scatter, const unsigned char* txt_pad,
                            const unsigned char* img_pad, int B, int S, int M, int N, int tl, const float* C_save,
                            const float* T_save, const float* ddist, void* dseq, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Optimiser on flat arenas: optim/adamw.py:40-103 (AdamW.step), pretrain.py:610 (clip_grad_norm_),
 * optimizer.zero_grad(), and the fp32 -> bf16 weight shadow refresh (replaces apex amp O2 master
 * weights, pretrain.py:463-465).  A "chunk" is a <= 8192-element slice of one parameter tensor.
 */
typedef struct { long long offset; int n; int tensor; } uc2_opt_chunk;
typedef struct {
    float lr[8]; float weight_decay[8];   /* per param group */
    float beta1, beta2, eps;
    int correct_bias;
    int global_step;                      /* 1-based count of optimizer steps */
    float max_grad_norm;                  /* <= 0: no clipping */
    int zero_grad;                        /* also clear the gradient arena */
} uc2_adamw_hyper;

UC2_API int uc2_cast_f32_bf16(const float* src, void* dst_bf16, long long n, void* stream);
/* *out (double, device) = sum of squares of the gradients of tensors with act_step >= 0 */
UC2_API int uc2_grad_sqnorm(const float* grad, const uc2_opt_chunk* chunks, int n_chunks, const int* act_step,
                            double* out, void* stream);
/* act_step[t]: global step at which tensor t first had a gradient (-1: never -> skipped, as p.grad is None);
 * group_of[t]: param group; sqnorm: device scalar from uc2_grad_sqnorm (may be NULL when not clipping). */
UC2_API int uc2_adamw_step(float* param, float* grad, float* exp_avg, float* exp_avg_sq, void* shadow_bf16,
                           const uc2_opt_chunk* chunks, int n_chunks, const int* act_step, const int* group_of,
                           const uc2_adamw_hyper* hyper, const double* sqnorm, void* stream);

/* Deferred ("lazy") AdamW for one row-sparse table -- the 250 002 x 768 word-embedding table, 69 % of all parameters.
 * A step whose only table gradient comes from the embedding lookup touches <= B*T rows, yet AdamW changes EVERY row
 * (moment decay + decoupled weight decay: optim/adamw.py:77-101), 34 bytes per element of HBM traffic for the eager
 * kernel.  Here a row's update is postponed until the row is needed -- it gets a gradient (uc2_adamw_lazy_rows), the
 * forward reads it, or someone looks at the parameters (uc2_adamw_lazy_catchup) -- and then every postponed step is
 * replayed on it in order with the identical fp32 operations and the identical per-step scalars (kept in a device
 * ring `hist`, written by the same device code that the eager kernel uses to derive them), so parameters and moments
 * end up BIT-IDENTICAL to eager AdamW while the rows stay in registers over the replay.
 * row_step[r] = last optimizer step applied to row r; row_seen = scratch for duplicate ids (init -1). */
typedef struct {
    float* param; float* grad; float* exp_avg; float* exp_avg_sq;   /* arena bases */
    long long table_off;      /* element offset of the table in the arenas */
    int n_rows, width;        /* width must be a multiple of 4 */
    int* row_step; int* row_seen;
    float* hist; int hist_len;/* ring of (step_size, lr * weight_decay) pairs, entry t % hist_len belongs to step t */
    float beta1, beta2, eps;
    int decay_on;             /* the table's group has weight_decay > 0 (the eager kernel's `if (wd > 0)`) */
    void* shadow_bf16;        /* base of the bf16 shadow arena (may be NULL): rows are re-cast whenever they change */
} uc2_lazy_table;
/* Step `step` (1-based) for the rows in row_ids (duplicates allowed, any order): replay what each row missed, apply
 * this step with its gradient (times the clip coefficient derived from *sqnorm as uc2_adamw_step does), clear the
 * gradient row.  first_step = the global step at which the table first had a gradient (bias correction counts from it). */
UC2_API int uc2_adamw_lazy_rows(const uc2_lazy_table* t, const long long* row_ids, long long n_ids, int step,
                                int first_step, float lr, float weight_decay, int correct_bias, float max_grad_norm,
                                const double* sqnorm, void* stream);
/* Record step `step`'s scalars in the ring WITHOUT touching any row (a step with no row list on this rank). */
UC2_API int uc2_adamw_lazy_note(const uc2_lazy_table* t, int step, int first_step, float lr, float weight_decay,
                                int correct_bias, void* stream);
/* Bring rows up to date through step `upto` (row_ids == NULL: every row). */
UC2_API int uc2_adamw_lazy_catchup(const uc2_lazy_table* t, const long long* row_ids, long long n_ids, int upto,
                                   void* stream);
/* *out += the table's share of the squared gradient norm on a row-sparse step: the eager kernel's per-chunk partial
 * sums (8192-element slices counted from table_off) of exactly those chunks that hold a row of row_ids, each distinct
 * chunk once (mark = a value no earlier call used; row_seen is the scratch) -- every other chunk of the table is zero. */
UC2_API int uc2_grad_sqnorm_rows(const uc2_lazy_table* t, const long long* row_ids, long long n_ids, int mark,
                                 double* out, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Device-side batch assembly (the step before the path) for region features that are
 * resident in HBM as a ragged arena: image i owns arena rows [row0_i, row0_i + nbb_i).
 * uc2_pad_rows replaces pad_tensors data/data.py:360-373 (out[b, r, :] = arena row, zeros for r >= nbb_b),
 * _mask_img_feat data/mrm.py:35-38 (zero_masked: masked rows written as zeros) and _get_feat_target 27-32 /
 * _get_targets 213-218 (targets[tgt_slot[b, r], :] = the unmasked row; tgt_slot = exclusive row-major scan of mask).
 * D is the row width: 2048 features, 7 box numbers, 1601 soft labels.  `out` may be NULL when only targets are wanted.
 */
typedef struct {
    const float* arena;          /* [rows_total, D] */
    int D;
    const long long* row0;       /* [B] first arena row of each batch item */
    const int* nbb;              /* [B] */
    const unsigned char* mask;   /* [B, R] 1 = masked region; may be NULL */
    const int* tgt_slot;         /* [B, R] row of `targets` for masked regions; may be NULL when targets is NULL */
    int B, R;
    int zero_masked;
} uc2_pad_args;
UC2_API int uc2_pad_rows(const uc2_pad_args* args, float* out, float* targets, void* stream);
/* Index side of a batch from the (txt_lens, num_bbs) vectors, T = max txt_len, R = max num_bb, S = packed length:
 *   attn_masks[b, j]   = j < tl_b + nbb_b                                   (data/itm.py:299-301 and siblings)
 *   gather_index[b, j] = j - tl_b + T on the image block, else j            (get_gather_index data/data.py:376-384)
 *   ot_scatter[b, j]   = j - tl_b + T for j >= tl_b, else j                 (_compute_ot_scatter data/itm.py:264-271)
 *   txt_pad[b, m] = m >= tl_b, img_pad[b, n] = n >= nbb_b                   (_compute_pad data/itm.py:274-278)
 *   img_mask_tgt[b, j] = img_mask[b, j - tl_b] on the image block, else 0   (_get_img_tgt_mask data/mrm.py:22-25)
 * Any output may be NULL. */
UC2_API int uc2_batch_index(const int* txt_lens, const int* num_bbs, const unsigned char* img_mask, int B, int T,
                            int R, int S, long long* attn_masks, long long* gather_index, long long* ot_scatter,
                            unsigned char* txt_pad, unsigned char* img_pad, unsigned char* img_mask_tgt, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* UC2_B200_H */
