// The BertLayer stack (model/layer.py:159-170 x num_hidden_layers, model/model.py:366-383 / 1031-1048)
// forward and backward, driven from C++ so that one ABI call enqueues every kernel of the stack.
// Per layer forward : QKV GEMM(+bias) -> flash attention -> O-proj GEMM(+bias+residual) -> LN
//                     -> FFN1 GEMM(+bias+erf-GELU, keeps pre-activation) -> FFN2 GEMM(+bias+residual) -> LN
// Per layer backward: LN bwd -> wgrad/dgrad(+dGELU) FFN2 -> wgrad/dgrad(+residual) FFN1 -> LN bwd
//                     -> wgrad/dgrad O-proj -> attention bwd -> wgrad/dgrad(+residual) QKV.
// The residual stream (LayerNorm inputs z1/z2 and the LayerNorm outputs that feed the next residual add) is
// fp32 in HBM; every GEMM operand is bf16.
#include "common.cuh"

namespace uc2 {
namespace {

constexpr int FF = 3072;
constexpr int QKV = 3 * HID;

struct G {
    const void* a; long long lda; bool a_mn;
    const void* b; long long ldb; bool b_mn;
    int M, N, K;
    const float* bias = nullptr;
    const void* residual = nullptr; long long ld_res = 0; bool res_f32 = false;
    const void* aux = nullptr; long long ld_aux = 0;
    int act = UC2_ACT_NONE;
    void* out_bf16 = nullptr; long long ld_out = 0;
    void* out_pre = nullptr; long long ld_pre = 0;
    float* out_f32 = nullptr; long long ld_f32 = 0;
    int accumulate = 0, split_k = 1;
    DropCfg drop = {0u, 0u, 1.f};
};

int run(cudaStream_t s, const G& g) {
    uc2_gemm_args a = {};          // every field this file does not set (the fused-CE ones) stays null
    a.a = g.a; a.lda = g.lda; a.a_mn = g.a_mn; a.b = g.b; a.ldb = g.ldb; a.b_mn = g.b_mn;
    a.M = g.M; a.N = g.N; a.K = g.K; a.bias = g.bias; a.residual = g.residual; a.ld_res = g.ld_res;
    a.aux = g.aux; a.ld_aux = g.ld_aux; a.act = g.act; a.out_bf16 = g.out_bf16; a.ld_out = g.ld_out;
    a.out_pre = g.out_pre; a.ld_pre = g.ld_pre; a.out_f32 = g.out_f32; a.ld_f32 = g.ld_f32;
    a.accumulate = g.accumulate; a.split_k = g.split_k; a.block_n = 0; a.residual_f32 = g.res_f32; a.ctas = 0;
    a.drop_key = g.drop.key; a.drop_thresh = g.drop.thresh; a.drop_scale = g.drop.scale; a.tail_split = 0;
    return uc2_gemm_bf16(&a, s);
}

// Y[M,N] = X[M,K] W[N,K]^T + bias
G linear_fwd(const void* x, const void* w, const float* bias, int M, int N, int K) {
    G g{x, K, false, w, K, false, M, N, K};
    g.bias = bias;
    return g;
}
// dX[M,K] = dY[M,N] W[N,K]
G linear_dgrad(const void* dy, const void* w, int M, int N, int K, void* dx) {
    G g{dy, N, false, w, K, true, M, K, N};
    g.out_bf16 = dx; g.ld_out = K;
    return g;
}
// dW[N,K] += dY[M,N]^T X[M,K]   (split-K over the token dimension, fp32 atomics)
G linear_wgrad(const void* dy, const void* x, int M, int N, int K, float* dw) {
    G g{dy, N, true, x, K, true, N, K, M};
    g.out_f32 = dw; g.ld_f32 = K; g.accumulate = 1; g.split_k = 0;
    return g;
}

size_t al256(size_t x) { return (x + 255) / 256 * 256; }

}  // namespace
}  // namespace uc2

using namespace uc2;

extern "C" UC2_API size_t uc2_encoder_fwd_workspace_bytes(int B, int S) {
    return 3 * al256((size_t)B * S * HID * 4);       // three rotating fp32 residual-stream buffers
}

extern "C" UC2_API size_t uc2_encoder_bwd_workspace_bytes(int B, int S) {
    const size_t M = (size_t)B * S;
    // 6 x [M,768] (two of them only used with dropout) + [M,3072] + [M,2304] bf16, + delta fp32 [B,12,S];
    // each region 256-byte aligned
    return 6 * al256(M * HID * 2) + al256(M * FF * 2) + al256(M * QKV * 2) + al256(M * 12 * 4);
}

extern "C" UC2_API int uc2_encoder_fwd(const void* x_in, const float* x_in_f32, const long long* attn_mask, int B,
                                       int S, int n_layers, const uc2_layer_weights* w, const uc2_layer_acts* acts,
                                       int save_for_bwd, void* workspace, size_t workspace_bytes, void* stream) {
    return uc2_encoder_fwd_dropout(x_in, x_in_f32, attn_mask, B, S, n_layers, w, acts, save_for_bwd, nullptr, workspace,
                                   workspace_bytes, stream);
}

extern "C" UC2_API int uc2_encoder_fwd_dropout(const void* x_in, const float* x_in_f32, const long long* attn_mask,
                                               int B, int S, int n_layers, const uc2_layer_weights* w,
                                               const uc2_layer_acts* acts, int save_for_bwd, const uc2_dropout* drop,
                                               void* workspace, size_t workspace_bytes, void* stream) {
    if (int rc = require_sm100()) return rc;
    const uc2_dropout nodrop = {0u, 1.f, 0u, 1.f};
    const uc2_dropout& D = drop ? *drop : nodrop;
    UC2_REQUIRE(x_in && x_in_f32 && attn_mask && w && acts && workspace && n_layers > 0 && B > 0 && S > 0, UC2_ERR_ARG,
                "encoder_fwd: bad args");
    UC2_REQUIRE(workspace_bytes >= uc2_encoder_fwd_workspace_bytes(B, S), UC2_ERR_ARG, "encoder_fwd: workspace too small");
    UC2_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, UC2_ERR_ARG, "encoder_fwd: workspace alignment");
    cudaStream_t s = (cudaStream_t)stream;
    const int M = B * S;
    float* res[3];
    for (int i = 0; i < 3; ++i)
        res[i] = reinterpret_cast<float*>(static_cast<uint8_t*>(workspace) + i * al256((size_t)M * HID * 4));
    const void* x = x_in;            // bf16 layer input (GEMM operand)
    const float* xr = x_in_f32;      // fp32 layer input (residual)
    int rot = 0;
    for (int l = 0; l < n_layers; ++l) {
        const uc2_layer_weights& W = w[l];
        const uc2_layer_acts& A = acts[l];
        UC2_REQUIRE(A.qkv && A.ctx && A.lse && A.z1 && A.h1 && A.g && A.z2 && A.out, UC2_ERR_ARG,
                    "encoder_fwd: layer %d activation buffers missing", l);
        float* h1r = res[rot];
        float* outr = res[(rot + 1) % 3];
        {   // Q, K, V in one GEMM
            G g = linear_fwd(x, W.w_qkv, W.b_qkv, M, QKV, HID);
            g.out_bf16 = A.qkv; g.ld_out = QKV;
            if (int rc = run(s, g)) return rc;
        }
        if (int rc = uc2_attention_fwd_dropout(A.qkv, attn_mask, A.ctx, A.lse, B, S, A.key_attn, D.attn_thresh,
                                               D.attn_scale, stream))
            return rc;
        {   // attention output projection + bias (+ dropout) + residual -> fp32 LayerNorm input
            G g = linear_fwd(A.ctx, W.w_o, W.b_o, M, HID, HID);
            g.drop = DropCfg{A.key_out1, D.hidden_thresh, D.hidden_scale};
            g.residual = xr; g.ld_res = HID; g.res_f32 = true;
            g.out_f32 = static_cast<float*>(A.z1); g.ld_f32 = HID;
            if (int rc = run(s, g)) return rc;
        }
        if (int rc = uc2_layernorm_fwd(A.z1, 1, W.ln1_w, W.ln1_b, 1e-12f, A.h1, h1r, M, stream)) return rc;
        {   // FFN1 + bias + erf-GELU (pre-activation kept for backward)
            G g = linear_fwd(A.h1, W.w_ffn1, W.b_ffn1, M, FF, HID);
            g.act = UC2_ACT_GELU; g.out_bf16 = A.g; g.ld_out = FF;
            if (save_for_bwd) { g.out_pre = A.u; g.ld_pre = FF; }
            if (int rc = run(s, g)) return rc;
        }
        {   // FFN2 + bias (+ dropout) + residual
            G g = linear_fwd(A.g, W.w_ffn2, W.b_ffn2, M, HID, FF);
            g.drop = DropCfg{A.key_out2, D.hidden_thresh, D.hidden_scale};
            g.residual = h1r; g.ld_res = HID; g.res_f32 = true;
            g.out_f32 = static_cast<float*>(A.z2); g.ld_f32 = HID;
            if (int rc = run(s, g)) return rc;
        }
        if (int rc = uc2_layernorm_fwd(A.z2, 1, W.ln2_w, W.ln2_b, 1e-12f, A.out, outr, M, stream)) return rc;
        x = A.out;
        xr = outr;
        rot = (rot + 2) % 3;         // h1r is dead; outr must survive the next layer's O-proj
    }
    return UC2_OK;
}

extern "C" UC2_API int uc2_encoder_bwd(const void* x_in, const long long* attn_mask, int B, int S, int n_layers,
                                       const uc2_layer_weights* w, const uc2_layer_acts* acts,
                                       const uc2_layer_grads* grads, const void* dout, void* dx_in, void* workspace,
                                       size_t workspace_bytes, void* stream) {
    return uc2_encoder_bwd_dropout(x_in, attn_mask, B, S, n_layers, w, acts, grads, dout, dx_in, nullptr, workspace,
                                   workspace_bytes, stream);
}

extern "C" UC2_API int uc2_encoder_bwd_dropout(const void* x_in, const long long* attn_mask, int B, int S, int n_layers,
                                               const uc2_layer_weights* w, const uc2_layer_acts* acts,
                                               const uc2_layer_grads* grads, const void* dout, void* dx_in,
                                               const uc2_dropout* drop, void* workspace, size_t workspace_bytes,
                                               void* stream) {
    if (int rc = require_sm100()) return rc;
    const uc2_dropout nodrop = {0u, 1.f, 0u, 1.f};
    const uc2_dropout& D = drop ? *drop : nodrop;
    const bool hd = D.hidden_thresh != 0;
    UC2_REQUIRE(x_in && attn_mask && w && acts && grads && dout && dx_in && workspace && n_layers > 0, UC2_ERR_ARG,
                "encoder_bwd: bad args");
    UC2_REQUIRE(workspace_bytes >= uc2_encoder_bwd_workspace_bytes(B, S), UC2_ERR_ARG,
                "encoder_bwd: workspace too small (%zu < %zu)", workspace_bytes, uc2_encoder_bwd_workspace_bytes(B, S));
    UC2_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, UC2_ERR_ARG, "encoder_bwd: workspace alignment");
    cudaStream_t s = (cudaStream_t)stream;
    const int M = B * S;
    uint8_t* p = static_cast<uint8_t*>(workspace);
    void* ws768[6];
    for (int i = 0; i < 6; ++i) { ws768[i] = p; p += al256((size_t)M * HID * 2); }
    void* du = p; p += al256((size_t)M * FF * 2);
    void* dqkv = p; p += al256((size_t)M * QKV * 2);
    float* delta = reinterpret_cast<float*>(p);

    const void* d_cur = dout;       // gradient w.r.t. the current layer's output
    int flip = 0;                   // dz2 (and then d_next, once dz2 is dead) alternate between slots 0 and 1,
                                    // so the d_cur read by a layer never aliases the dz2 it writes
    for (int l = n_layers - 1; l >= 0; --l) {
        const uc2_layer_weights& W = w[l];
        const uc2_layer_acts& A = acts[l];
        const uc2_layer_grads& Gr = grads[l];
        const void* x_l = l == 0 ? x_in : acts[l - 1].out;
        void* dz2 = ws768[flip];
        void* dh1 = ws768[2];                   // later reused for dctx
        void* dz1 = ws768[3];
        // with hidden dropout the dense branches see the masked, rescaled gradients; the residual adds the plain ones
        void* dz2m = hd ? ws768[4] : dz2;
        void* dz1m = hd ? ws768[5] : dz1;
        void* d_next = l == 0 ? dx_in : dz2;    // dz2 is last read by the FFN1 dgrad, long before this is written
        UC2_REQUIRE(A.u, UC2_ERR_ARG, "encoder_bwd: layer %d has no saved pre-GELU activation", l);
        // output LayerNorm + FFN2
        if (int rc = uc2_layernorm_bwd_dropout(A.z2, 1, d_cur, W.ln2_w, 1e-12f, dz2, Gr.ln2_w, Gr.ln2_b, Gr.b_ffn2, M,
                                               hd ? dz2m : nullptr, A.key_out2, D.hidden_thresh, D.hidden_scale, stream))
            return rc;
        if (int rc = run(s, linear_wgrad(dz2m, A.g, M, HID, FF, Gr.w_ffn2))) return rc;
        {
            G g = linear_dgrad(dz2m, W.w_ffn2, M, HID, FF, du);
            g.aux = A.u; g.ld_aux = FF; g.act = UC2_ACT_DGELU;
            if (int rc = run(s, g)) return rc;
        }
        // FFN1
        if (int rc = uc2_colsum_bf16(du, FF, M, FF, Gr.b_ffn1, stream)) return rc;
        if (int rc = run(s, linear_wgrad(du, A.h1, M, FF, HID, Gr.w_ffn1))) return rc;
        {
            G g = linear_dgrad(du, W.w_ffn1, M, FF, HID, dh1);
            g.residual = dz2; g.ld_res = HID;
            if (int rc = run(s, g)) return rc;
        }
        // attention-output LayerNorm + projection
        if (int rc = uc2_layernorm_bwd_dropout(A.z1, 1, dh1, W.ln1_w, 1e-12f, dz1, Gr.ln1_w, Gr.ln1_b, Gr.b_o, M,
                                               hd ? dz1m : nullptr, A.key_out1, D.hidden_thresh, D.hidden_scale, stream))
            return rc;
        if (int rc = run(s, linear_wgrad(dz1m, A.ctx, M, HID, HID, Gr.w_o))) return rc;
        void* dctx = dh1;
        if (int rc = run(s, linear_dgrad(dz1m, W.w_o, M, HID, HID, dctx))) return rc;
        // attention + QKV projection
        if (int rc = uc2_attention_bwd_dropout(A.qkv, attn_mask, A.ctx, dctx, A.lse, delta, dqkv, B, S, A.key_attn,
                                               D.attn_thresh, D.attn_scale, stream))
            return rc;
        if (int rc = uc2_colsum_bf16(dqkv, QKV, M, QKV, Gr.b_qkv, stream)) return rc;
        if (int rc = run(s, linear_wgrad(dqkv, x_l, M, QKV, HID, Gr.w_qkv))) return rc;
        {
            G g = linear_dgrad(dqkv, W.w_qkv, M, QKV, HID, d_next);
            g.residual = dz1; g.ld_res = HID;
            if (int rc = run(s, g)) return rc;
        }
        d_cur = d_next;
        flip ^= 1;
    }
    return UC2_OK;
}
