"""ctypes binding of libuc2_b200.so (the C ABI declared in include/uc2_b200.h).

There is no CPU fallback: if the shared library is missing or a call fails, a
RuntimeError is raised with the library's own message.
"""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libuc2_b200.so")
_lib = None

ACT_NONE, ACT_GELU, ACT_DGELU, ACT_TANH = 0, 1, 2, 3
P, LL, I, F, SZ = C.c_void_p, C.c_longlong, C.c_int, C.c_float, C.c_size_t


class GemmArgs(C.Structure):
    _fields_ = [("a", P), ("lda", LL), ("a_mn", I), ("b", P), ("ldb", LL), ("b_mn", I),
                ("M", I), ("N", I), ("K", I), ("bias", P), ("residual", P), ("ld_res", LL),
                ("aux", P), ("ld_aux", LL), ("act", I), ("out_bf16", P), ("ld_out", LL),
                ("out_pre", P), ("ld_pre", LL), ("out_f32", P), ("ld_f32", LL),
                ("accumulate", I), ("split_k", I), ("block_n", I), ("residual_f32", I), ("ctas", I),
                ("drop_key", C.c_uint), ("drop_thresh", C.c_uint), ("drop_scale", F), ("tail_split", I),
                ("ce_stats", P), ("ld_ce", LL), ("ce_labels", P), ("ce_tgt", P)]


class EmbedArgs(C.Structure):
    _fields_ = [("B", I), ("T", I), ("R", I), ("S", I), ("mode", I), ("hidden", I),
                ("input_ids", P), ("position_ids", P), ("position_rows", I), ("gather_index", P),
                ("word_pad_id", I), ("pos_pad_id", I),
                ("word_emb", P), ("pos_emb", P), ("type_emb", P), ("ln_w", P), ("ln_b", P),
                ("y_img", P), ("img_pos_feat", P), ("img_ln_w", P), ("img_ln_b", P),
                ("pos_w", P), ("pos_b", P), ("pos_ln_w", P), ("pos_ln_b", P),
                ("fin_ln_w", P), ("fin_ln_b", P), ("eps", F), ("vocab", I), ("max_pos", I),
                ("drop_key", C.c_uint), ("drop_thresh", C.c_uint), ("drop_scale", F)]


class EmbedGrads(C.Structure):
    _fields_ = [(n, P) for n in ("word_emb", "pos_emb", "type_emb", "ln_w", "ln_b", "img_ln_w", "img_ln_b",
                                 "pos_w", "pos_b", "pos_ln_w", "pos_ln_b", "fin_ln_w", "fin_ln_b", "dy_img")]


LAYER_FIELDS = ("w_qkv", "b_qkv", "w_o", "b_o", "ln1_w", "ln1_b", "w_ffn1", "b_ffn1", "w_ffn2", "b_ffn2",
                "ln2_w", "ln2_b")
ACT_FIELDS = ("qkv", "ctx", "lse", "z1", "h1", "u", "g", "z2", "out")


class LayerWeights(C.Structure):
    _fields_ = [(n, P) for n in LAYER_FIELDS]


class LayerGrads(C.Structure):
    _fields_ = [(n, P) for n in LAYER_FIELDS]


class LayerActs(C.Structure):
    _fields_ = [(n, P) for n in ACT_FIELDS] + [("key_attn", C.c_uint), ("key_out1", C.c_uint), ("key_out2", C.c_uint)]


class Dropout(C.Structure):
    _fields_ = [("attn_thresh", C.c_uint), ("attn_scale", F), ("hidden_thresh", C.c_uint), ("hidden_scale", F)]


class PadArgs(C.Structure):
    _fields_ = [("arena", P), ("D", I), ("row0", P), ("nbb", P), ("mask", P), ("tgt_slot", P), ("B", I), ("R", I),
                ("zero_masked", I)]


class OptChunk(C.Structure):
    _fields_ = [("offset", LL), ("n", I), ("tensor", I)]


class LazyTable(C.Structure):
    _fields_ = [("param", P), ("grad", P), ("exp_avg", P), ("exp_avg_sq", P), ("table_off", LL), ("n_rows", I),
                ("width", I), ("row_step", P), ("row_seen", P), ("hist", P), ("hist_len", I),
                ("beta1", F), ("beta2", F), ("eps", F), ("decay_on", I), ("shadow_bf16", P)]


class AdamwHyper(C.Structure):
    _fields_ = [("lr", F * 8), ("weight_decay", F * 8), ("beta1", F), ("beta2", F), ("eps", F),
                ("correct_bias", I), ("global_step", I), ("max_grad_norm", F), ("zero_grad", I)]


_SIGS = {
    "uc2_gemm_bf16": [C.POINTER(GemmArgs), P],
    "uc2_img_prep": [P, P, P, P, LL, I, P],
    "uc2_embed_pack_fwd": [C.POINTER(EmbedArgs), P, P, P],
    "uc2_embed_pack_bwd": [C.POINTER(EmbedArgs), P, C.POINTER(EmbedGrads), P],
    "uc2_img_grad_finish": [P, P, P, P, P, LL, P],
    "uc2_vecmat_acc": [P, P, P, I, I, P],
    "uc2_layernorm_fwd": [P, I, P, P, F, P, P, LL, P],
    "uc2_layernorm_bwd": [P, I, P, P, F, P, P, P, P, LL, P],
    "uc2_layernorm_bwd_dropout": [P, I, P, P, F, P, P, P, P, LL, P, C.c_uint, C.c_uint, F, P],
    "uc2_attention_fwd_dropout": [P, P, P, P, I, I, C.c_uint, C.c_uint, F, P],
    "uc2_attention_bwd_dropout": [P, P, P, P, P, P, P, I, I, C.c_uint, C.c_uint, F, P],
    "uc2_attention_fwd_tc": [P, P, P, P, I, I, C.c_uint, C.c_uint, F, P],
    "uc2_attention_bwd_tc": [P, P, P, P, P, P, I, I, C.c_uint, C.c_uint, F, P],
    "uc2_encoder_fwd_dropout": [P, P, P, I, I, I, C.POINTER(LayerWeights), C.POINTER(LayerActs), I, C.POINTER(Dropout),
                                P, SZ, P],
    "uc2_encoder_bwd_dropout": [P, P, I, I, I, C.POINTER(LayerWeights), C.POINTER(LayerActs), C.POINTER(LayerGrads),
                                P, P, C.POINTER(Dropout), P, SZ, P],
    "uc2_colsum_bf16": [P, LL, LL, I, P, P],
    "uc2_attention_fwd": [P, P, P, P, I, I, P],
    "uc2_attention_bwd": [P, P, P, P, P, P, P, I, I, P],
    "uc2_encoder_fwd": [P, P, P, I, I, I, C.POINTER(LayerWeights), C.POINTER(LayerActs), I, P, SZ, P],
    "uc2_encoder_bwd": [P, P, I, I, I, C.POINTER(LayerWeights), C.POINTER(LayerActs), C.POINTER(LayerGrads),
                        P, P, P, SZ, P],
    "uc2_narrow_linear_fwd": [P, LL, P, P, P, I, I, I, P],
    "uc2_narrow_linear_bwd": [P, LL, P, P, P, LL, P, P, I, I, I, P],
    "uc2_tanh_bwd": [P, P, P, LL, P],
    "uc2_rank_loss_fwd": [P, P, I, I, F, P],
    "uc2_rank_loss_bwd": [P, P, P, I, I, F, P],
    "uc2_softmax_loss": [P, LL, LL, I, I, P, LL, P, P, P, P, P, P],
    "uc2_mse": [P, P, P, P, P, LL, P],
    "uc2_ce_loss_fwd": [P, LL, LL, I, P, LL, P, P, P],
    "uc2_ce_stats_reduce": [P, LL, I, LL, P, P, LL, P, P, P, P],
    "uc2_ce_bwd_inplace_bf16": [P, LL, LL, I, P, LL, P, P, P],
    "uc2_ce_loss_bwd_bf16": [P, LL, LL, I, P, LL, P, P, P, LL, P],
    "uc2_mask_scan": [P, LL, P, P, I, P],
    "uc2_gather_rows": [P, P, P, I, I, P, I, P],
    "uc2_scatter_rows_add": [P, P, P, I, I, P, I, P],
    "uc2_dgelu_bf16": [P, P, P, LL, P],
    "uc2_f32_to_bf16_2d": [P, LL, P, LL, LL, I, P],
    "uc2_ot_ipot_fwd": [P, P, P, P, I, I, I, I, I, F, I, I, P, P, P, P],
    "uc2_ot_ipot_bwd": [P, P, P, P, I, I, I, I, I, P, P, P, P, P],
    "uc2_pad_rows": [C.POINTER(PadArgs), P, P, P],
    "uc2_batch_index": [P, P, P, I, I, I, I, P, P, P, P, P, P, P],
    "uc2_profile_enable": [I],
    "uc2_profile_collect": [P, P, P, I],
    "uc2_cast_f32_bf16": [P, P, LL, P],
    "uc2_grad_sqnorm": [P, P, I, P, P, P],
    "uc2_adamw_step": [P, P, P, P, P, P, I, P, P, C.POINTER(AdamwHyper), P, P],
    "uc2_adamw_lazy_rows": [C.POINTER(LazyTable), P, LL, I, I, F, F, I, F, P, P],
    "uc2_adamw_lazy_note": [C.POINTER(LazyTable), I, I, F, F, I, P],
    "uc2_adamw_lazy_catchup": [C.POINTER(LazyTable), P, LL, I, P],
    "uc2_grad_sqnorm_rows": [C.POINTER(LazyTable), P, LL, I, P, P],
}
EXPORTS = sorted(list(_SIGS) + ["uc2_last_error", "uc2_version", "uc2_launch_count", "uc2_attention_tc_enable", "uc2_reserve_sms", "uc2_gemm_sched_dynamic", "uc2_debug_occupy_sms",
                                "uc2_encoder_bwd_workspace_bytes", "uc2_encoder_fwd_workspace_bytes"])


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"uc2_b200: {LIB_PATH} not found. Build it with `python -m uc2_b200.build` "
                "(there is no CPU / PyTorch fallback for this path).")
        L = C.CDLL(LIB_PATH)
        L.uc2_last_error.restype = C.c_char_p
        L.uc2_launch_count.restype = C.c_longlong
        L.uc2_encoder_bwd_workspace_bytes.restype = SZ
        L.uc2_encoder_bwd_workspace_bytes.argtypes = [I, I]
        L.uc2_encoder_fwd_workspace_bytes.restype = SZ
        L.uc2_encoder_fwd_workspace_bytes.argtypes = [I, I]
        L.uc2_attention_tc_enable.restype = I
        L.uc2_attention_tc_enable.argtypes = [I]
        L.uc2_gemm_sched_dynamic.restype = I
        L.uc2_gemm_sched_dynamic.argtypes = [I]
        L.uc2_reserve_sms.restype = I
        L.uc2_reserve_sms.argtypes = [I]
        L.uc2_debug_occupy_sms.restype = I
        L.uc2_debug_occupy_sms.argtypes = [I, LL, P]
        for name, sig in _SIGS.items():
            fn = getattr(L, name)
            fn.argtypes = sig
            fn.restype = I
        _lib = L
    return _lib


def check(rc, what=""):
    if rc != 0:
        msg = lib().uc2_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"uc2_b200 {what} failed (code {rc}): {msg}")


def stream():
    return torch.cuda.current_stream().cuda_stream


def ptr(t):
    return None if t is None else t.data_ptr()


def call(name, *args):
    check(getattr(lib(), name)(*args), name)


def attention_tc_enabled():
    """Whether the public attention entry points route to the tcgen05 kernels (the default; UC2_ATTN_TCGEN05=0 or
    uc2_attention_tc_enable(0) switch back to the mma.sync kernels of csrc/attention.cu)."""
    prev = lib().uc2_attention_tc_enable(0)
    lib().uc2_attention_tc_enable(prev)
    return bool(prev)


def attention_kernels(S):
    """Which attention kernels the public entry points run for a packed length S (for the bench line)."""
    if not attention_tc_enabled() or S > 256:
        return "mma.sync (attention.cu)"
    return "tcgen05/TMEM forward + backward (attention_tc.cu)"


def launch_count():
    return int(lib().uc2_launch_count())


def gemm(a, b, M, N, K, *, a_mn=False, b_mn=False, lda=None, ldb=None, bias=None, residual=None, aux=None,
         act=ACT_NONE, out_bf16=None, out_pre=None, out_f32=None, accumulate=False, split_k=1, block_n=0,
         ld_out=None, ld_res=None, ctas=0, drop=None, tail_split=0, ce=None):
    """D[M,N] = A[M,K] @ B[N,K]^T with the fused epilogue described in include/uc2_b200.h."""
    g = GemmArgs()
    g.a, g.lda, g.a_mn = a.data_ptr(), (lda if lda is not None else a.stride(0)), int(a_mn)
    g.b, g.ldb, g.b_mn = b.data_ptr(), (ldb if ldb is not None else b.stride(0)), int(b_mn)
    g.M, g.N, g.K = M, N, K
    g.bias = bias.data_ptr() if bias is not None else None
    if residual is not None:
        g.residual, g.ld_res = residual.data_ptr(), (ld_res if ld_res is not None else residual.stride(0))
        g.residual_f32 = int(residual.dtype == torch.float32)
    if aux is not None:
        g.aux, g.ld_aux = aux.data_ptr(), aux.stride(0)
    g.act = act
    if out_bf16 is not None:
        g.out_bf16, g.ld_out = out_bf16.data_ptr(), (ld_out if ld_out is not None else out_bf16.stride(0))
    if out_pre is not None:
        g.out_pre, g.ld_pre = out_pre.data_ptr(), out_pre.stride(0)
    if out_f32 is not None:
        g.out_f32, g.ld_f32 = out_f32.data_ptr(), out_f32.stride(0)
    g.accumulate, g.split_k, g.block_n, g.ctas = int(accumulate), split_k, block_n, ctas
    g.tail_split = tail_split
    if ce is not None:                     # (stats [n_chunks, ld_ce, 2] fp32, labels int64 [M], target logits fp32 [M])
        stats, labels, tgt = ce
        g.ce_stats, g.ld_ce, g.ce_labels, g.ce_tgt = stats.data_ptr(), stats.stride(0) // 2, labels.data_ptr(), tgt.data_ptr()
    if drop is not None:
        g.drop_key, g.drop_thresh, g.drop_scale = drop
    check(lib().uc2_gemm_bf16(C.byref(g), stream()), "gemm")
