"""torch.autograd wrappers over the C ABI.  Every Function computes its activation gradient and
adds its parameter gradients straight into the gradient arena (``arena.g(name)``, which IS ``p.grad``),
so autograd only carries activations between the fused blocks.
"""
import ctypes as C

import torch

from . import _lib
from . import dropout as DO
from ._lib import call, ptr, stream

BF16, F32 = torch.bfloat16, torch.float32
ctypes_size = C.sizeof


def _empty(shape, dtype, like):
    return torch.empty(shape, dtype=dtype, device=like.device)


# --------------------------------------------------------------------------------------------------
# encoder (embeddings + pack + BertLayer stack)
# --------------------------------------------------------------------------------------------------
class EncoderState(object):
    pass


def encoder_forward(enc, arena, input_ids, position_ids, img_feat, img_pos_feat, attention_mask, gather_index,
                    img_masks, save, embed_only=False, keep_all=False, dropout=None):
    """Runs K1-K10.  Returns (x0, [layer outputs], state).
    dropout: None, or (p_hidden, p_attn, seed, counter) -- the five nn.Dropout sites of the reference
    (model.py:334, 363; layer.py:94, 113, 154) as counter-based masks regenerated in backward (dropout.py)."""
    cfg, fam = enc.config, enc.family
    pre = enc.prefix
    dev = attention_mask.device
    st = EncoderState()
    am = attention_mask.to(torch.long).contiguous()
    B, S = am.shape
    has_txt, has_img = input_ids is not None, img_feat is not None
    mode = 0 if (has_txt and has_img) else (1 if has_txt else 2)
    T = input_ids.size(1) if has_txt else 0
    R = img_feat.size(1) if has_img else 0
    if mode == 0 and gather_index is None:
        raise ValueError("gather_index is required when both text and image inputs are given")
    a = _lib.EmbedArgs()
    a.B, a.T, a.R, a.S, a.mode, a.hidden = B, T, R, S, mode, cfg.hidden_size
    a.word_pad_id, a.pos_pad_id = fam.word_pad, fam.pos_pad
    a.eps = fam.emb_eps(cfg)
    a.vocab, a.max_pos = cfg.vocab_size, cfg.max_position_embeddings
    a.drop_scale = 1.0
    drop = None
    if dropout is not None:
        p_h, p_a, seed, counter = dropout
        a.drop_key, a.drop_thresh = DO.site_key(seed, counter, 255, DO.SITE_EMB), DO.thresh_of(p_h)
        a.drop_scale = DO.scale_of(p_h)
        drop = _lib.Dropout(DO.thresh_of(p_a), DO.scale_of(p_a), DO.thresh_of(p_h), DO.scale_of(p_h))
    a.type_emb = arena.mp(pre + fam.type_emb)
    keep = []
    if has_txt:
        ids = input_ids.to(torch.long).contiguous()
        keep.append(ids)
        a.input_ids = ids.data_ptr()
        if position_ids is not None:
            pos = position_ids.to(torch.long).contiguous()
            if pos.dim() == 1:
                pos = pos.unsqueeze(0)
            keep.append(pos)
            a.position_ids, a.position_rows = pos.data_ptr(), pos.size(0)
        elif not fam.derive_positions:
            raise ValueError("position_ids is required for the Uniter family")
        lazy = getattr(arena, "lazy", None)
        if lazy is not None:
            lazy.catch_up(ids)              # deferred AdamW (optim.LazyRows): the rows about to be read become current
        a.word_emb = arena.mp(pre + "embeddings.word_embeddings.weight")
        a.pos_emb = arena.mp(pre + "embeddings.position_embeddings.weight")
        a.ln_w = arena.mp(pre + "embeddings.LayerNorm.weight")
        a.ln_b = arena.mp(pre + "embeddings.LayerNorm.bias")
    feat_bf16 = y_img = masks_u8 = None
    if has_img:
        ip = pre + "img_embeddings."
        feat = img_feat.to(F32).contiguous()
        posf = img_pos_feat.to(F32).contiguous()
        keep.append(posf)
        if img_masks is not None:
            masks_u8 = img_masks.to(torch.uint8).contiguous()
            with torch.no_grad():
                arena.m(ip + "mask_embedding.weight")[0].zero_()      # model.py:354
        feat_bf16 = _empty((B * R, feat.size(-1)), BF16, feat)
        call("uc2_img_prep", feat.data_ptr(), ptr(masks_u8), arena.mp(ip + "mask_embedding.weight") + 4 * feat.size(-1),
             feat_bf16.data_ptr(), B * R, feat.size(-1), stream())
        y_img = _empty((B * R, cfg.hidden_size), F32, feat)
        _lib.gemm(feat_bf16, arena.s(ip + "img_linear.weight"), B * R, cfg.hidden_size, feat.size(-1),
                  bias=arena.m(ip + "img_linear.bias"), out_f32=y_img)
        a.y_img, a.img_pos_feat = y_img.data_ptr(), posf.data_ptr()
        a.img_ln_w, a.img_ln_b = arena.mp(ip + "img_layer_norm.weight"), arena.mp(ip + "img_layer_norm.bias")
        a.pos_w, a.pos_b = arena.mp(ip + "pos_linear.weight"), arena.mp(ip + "pos_linear.bias")
        a.pos_ln_w, a.pos_ln_b = arena.mp(ip + "pos_layer_norm.weight"), arena.mp(ip + "pos_layer_norm.bias")
        a.fin_ln_w, a.fin_ln_b = arena.mp(ip + "LayerNorm.weight"), arena.mp(ip + "LayerNorm.bias")
    if mode == 0:
        gi = gather_index.to(torch.long).contiguous()
        keep.append(gi)
        a.gather_index = gi.data_ptr()
    M = B * S
    x0 = torch.empty((B, S, cfg.hidden_size), dtype=BF16, device=dev)
    x0_f32 = None if embed_only else torch.empty((B, S, cfg.hidden_size), dtype=F32, device=dev)
    call("uc2_embed_pack_fwd", C.byref(a), x0.data_ptr(), ptr(x0_f32), stream())
    if embed_only:
        return x0, [], None

    L = cfg.num_hidden_layers
    W = enc.layer_weight_structs(arena)
    acts = (_lib.LayerActs * L)()
    bufs = []
    shared = None
    outs = []
    for l in range(L):
        if save or shared is None:
            b = dict(qkv=torch.empty((M, 2304), dtype=BF16, device=dev), ctx=torch.empty((M, 768), dtype=BF16, device=dev),
                     lse=torch.empty((B, 12, S), dtype=F32, device=dev), z1=torch.empty((M, 768), dtype=F32, device=dev),
                     h1=torch.empty((M, 768), dtype=BF16, device=dev),
                     u=torch.empty((M, 3072), dtype=BF16, device=dev) if save else None,
                     g=torch.empty((M, 3072), dtype=BF16, device=dev), z2=torch.empty((M, 768), dtype=F32, device=dev))
            shared = b
        else:
            b = dict(shared)
        b["out"] = torch.empty((B, S, 768), dtype=BF16, device=dev)
        bufs.append(b)
        outs.append(b["out"])
        for f in _lib.ACT_FIELDS:
            setattr(acts[l], f, ptr(b[f]))
        if dropout is not None:
            acts[l].key_attn = DO.site_key(seed, counter, l, DO.SITE_ATTN)
            acts[l].key_out1 = DO.site_key(seed, counter, l, DO.SITE_OUT1)
            acts[l].key_out2 = DO.site_key(seed, counter, l, DO.SITE_OUT2)
    fws_bytes = int(_lib.lib().uc2_encoder_fwd_workspace_bytes(B, S))
    fws = torch.empty(fws_bytes, dtype=torch.uint8, device=dev)
    call("uc2_encoder_fwd_dropout", x0.data_ptr(), x0_f32.data_ptr(), am.data_ptr(), B, S, L, W, acts, int(save),
         C.byref(drop) if drop is not None else None, fws.data_ptr(), fws_bytes, stream())
    st.__dict__.update(args=a, keep=keep, am=am, B=B, S=S, T=T, R=R, mode=mode, x0=x0, acts=acts, bufs=bufs,
                       feat_bf16=feat_bf16, y_img=y_img, masks_u8=masks_u8, W=W, drop=drop)
    return x0, outs, st


def encoder_backward(enc, arena, st, dout):
    """Backward of encoder_forward given d(last layer output) [B,S,768] bf16."""
    cfg, fam, pre = enc.config, enc.family, enc.prefix
    B, S, R = st.B, st.S, st.R
    M = B * S
    L = cfg.num_hidden_layers
    dout = dout.to(BF16).contiguous()
    G = enc.layer_grad_structs(arena)
    ws_bytes = int(_lib.lib().uc2_encoder_bwd_workspace_bytes(B, S))
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dout.device)
    dx0 = torch.empty((M, 768), dtype=BF16, device=dout.device)
    dptr = C.byref(st.drop) if st.drop is not None else None
    sync = getattr(arena, "grad_sync", None)
    if sync is None or not sync.enabled:
        call("uc2_encoder_bwd_dropout", st.x0.data_ptr(), st.am.data_ptr(), B, S, L, st.W, st.acts, G, dout.data_ptr(),
             dx0.data_ptr(), dptr, ws.data_ptr(), ws_bytes, stream())
    else:
        # Data parallel: run the stack in segments of layers (top first) and hand each finished slice of the
        # gradient arena to the NCCL stream while the next segment computes.
        q0 = lambda l: arena.offset[pre + f"encoder.layer.{l}.attention.self.query.weight"]
        last = pre + f"encoder.layer.{L - 1}.output.LayerNorm.bias"
        layers_end = arena.offset[last] + arena.numel[last]
        sync.ready(layers_end, arena.total)                 # pooler + heads: complete before this backward runs
        seg = max(1, getattr(sync, "layers_per_segment", 3))
        WS, AS, GS = ctypes_size(_lib.LayerWeights), ctypes_size(_lib.LayerActs), ctypes_size(_lib.LayerGrads)
        d_hi = dout
        hi = L
        while hi > 0:
            lo = max(0, hi - seg)
            x_in = st.x0 if lo == 0 else st.bufs[lo - 1]["out"]
            d_lo = dx0 if lo == 0 else torch.empty((M, 768), dtype=BF16, device=dout.device)
            call("uc2_encoder_bwd_dropout", x_in.data_ptr(), st.am.data_ptr(), B, S, hi - lo,
                 C.cast(C.byref(st.W, lo * WS), C.POINTER(_lib.LayerWeights)),
                 C.cast(C.byref(st.acts, lo * AS), C.POINTER(_lib.LayerActs)),
                 C.cast(C.byref(G, lo * GS), C.POINTER(_lib.LayerGrads)),
                 d_hi.data_ptr(), d_lo.data_ptr(), dptr, ws.data_ptr(), ws_bytes, stream())
            end = layers_end if hi == L else q0(hi)
            sync.ready(q0(lo), end)
            d_hi, hi = d_lo, lo
    arena.touch(*enc.layer_param_names())
    # embeddings
    g = _lib.EmbedGrads()
    names = []
    dense_sent = st.mode != 2 and sync is not None and sync.enabled and getattr(arena, "word_dense_sent", False)
    if st.mode != 2:
        wtab = pre + "embeddings.word_embeddings.weight"
        if dense_sent:
            # the table's dense (decoder) gradient is already being averaged: the lookup's rows go to a side table
            side = getattr(arena, "word_side", None)
            if side is None or side.shape != arena.shape[wtab]:
                side = arena.word_side = torch.zeros(arena.shape[wtab], dtype=F32, device=dout.device)
            g.word_emb = side.data_ptr()
        else:
            g.word_emb = arena.gp(wtab)
        g.pos_emb = arena.gp(pre + "embeddings.position_embeddings.weight")
        g.ln_w = arena.gp(pre + "embeddings.LayerNorm.weight")
        g.ln_b = arena.gp(pre + "embeddings.LayerNorm.bias")
        names += [pre + "embeddings.word_embeddings.weight", pre + "embeddings.position_embeddings.weight",
                  pre + "embeddings.LayerNorm.weight", pre + "embeddings.LayerNorm.bias"]
    g.type_emb = arena.gp(pre + fam.type_emb)
    names.append(pre + fam.type_emb)
    dy_img = None
    if st.mode != 1:
        ip = pre + "img_embeddings."
        dy_img = torch.zeros((B * R, 768), dtype=F32, device=dout.device)
        g.dy_img = dy_img.data_ptr()
        for f, n in (("img_ln_w", "img_layer_norm.weight"), ("img_ln_b", "img_layer_norm.bias"),
                     ("pos_w", "pos_linear.weight"), ("pos_b", "pos_linear.bias"),
                     ("pos_ln_w", "pos_layer_norm.weight"), ("pos_ln_b", "pos_layer_norm.bias"),
                     ("fin_ln_w", "LayerNorm.weight"), ("fin_ln_b", "LayerNorm.bias")):
            setattr(g, f, arena.gp(ip + n))
            names.append(ip + n)
    call("uc2_embed_pack_bwd", C.byref(st.args), dx0.data_ptr(), C.byref(g), stream())
    if st.mode != 2:
        # rows of the vocabulary table that now carry gradient, for the deferred AdamW (optim.LazyRows); a second
        # backward pass into the same optimizer step makes the list incomplete -> None = treat the table as dense
        arena.word_rows_n = getattr(arena, "word_rows_n", 0) + 1
        arena.word_rows = st.keep[0] if arena.word_rows_n == 1 else None
    if st.mode != 1:
        ip = pre + "img_embeddings."
        dy_bf16 = torch.empty((B * R, 768), dtype=BF16, device=dout.device)
        msum = torch.zeros(768, dtype=F32, device=dout.device) if st.masks_u8 is not None else None
        call("uc2_img_grad_finish", dy_img.data_ptr(), ptr(st.masks_u8), dy_bf16.data_ptr(),
             arena.gp(ip + "img_linear.bias"), ptr(msum), B * R, stream())
        D = st.feat_bf16.size(1)
        # dW_img[768, D] += dy^T feat
        _lib.gemm(dy_bf16, st.feat_bf16, 768, D, B * R, a_mn=True, b_mn=True, out_f32=arena.g(ip + "img_linear.weight"),
                  accumulate=True, split_k=0)
        names += [ip + "img_linear.weight", ip + "img_linear.bias"]
        if msum is not None:
            # d mask_embedding.weight[1] = (sum of masked rows of dy) @ img_linear.weight; row 0 is padding_idx
            call("uc2_vecmat_acc", msum.data_ptr(), arena.mp(ip + "img_linear.weight"),
                 arena.gp(ip + "mask_embedding.weight") + 4 * D, 768, D, stream())
            names.append(ip + "mask_embedding.weight")
    arena.touch(*names)
    if sync is not None and sync.enabled:
        q0_off = arena.offset[pre + "encoder.layer.0.attention.self.query.weight"]
        wname = pre + "embeddings.word_embeddings.weight"
        if dense_sent:
            sync.side_rows(arena.word_side, arena.g(wname), st.keep[0], pad_row=max(int(fam.word_pad), 0))
            arena.word_rows = None
            sync.ready(arena.numel[wname], q0_off)           # the table sits at the start of the arena
        elif (st.mode != 2 and arena.offset[wname] == 0 and getattr(sync, "allow_sparse", False)
                and not getattr(arena, "word_emb_dense", False)):
            # only token rows of the vocabulary table carry gradient: exchange those rows, not 768 MB of zeros
            w_end = arena.numel[wname]
            sync.sparse_rows_table(0, arena.shape[wname][0], arena.shape[wname][1], st.keep[0],
                                   pad_row=max(int(fam.word_pad), 0))
            if arena.word_rows is not None:
                arena.word_rows = sync.last_ids          # after the exchange: the rows of every rank
            sync.ready(w_end, q0_off)
        else:
            arena.word_rows = None                       # dense all-reduce: any rank's rows may carry gradient now
            sync.ready(0, q0_off)                                                        # embeddings: last


class EncoderFn(torch.autograd.Function):
    """(anchor) -> last hidden state; the other inputs are captured non-differentiably."""

    @staticmethod
    def forward(ctx, anchor, enc, arena, kw):
        x0, outs, st = encoder_forward(enc, arena, save=True, **kw)
        ctx.enc, ctx.arena, ctx.st = enc, arena, st
        return outs[-1]

    @staticmethod
    def backward(ctx, dout):
        encoder_backward(ctx.enc, ctx.arena, ctx.st, dout)
        ctx.st = None
        return None, None, None, None


# --------------------------------------------------------------------------------------------------
# dense / layernorm blocks on compacted rows (heads)
# --------------------------------------------------------------------------------------------------
class LinearFn(torch.autograd.Function):
    """y = act(x @ W^T + b).  x bf16 [n,K]; W is the bf16 shadow of parameter `wname`.
    transposed=True uses W^T (F.linear(h, weight.t(), bias), model.py:1155).  Output bf16, or fp32 when
    out_f32 (logits)."""

    @staticmethod
    def forward(ctx, x, arena, wname, bname, act, transposed, out_f32):
        lazy = getattr(arena, "lazy", None)
        if lazy is not None and wname == lazy.name:
            lazy.catch_up_all()             # the tied decoder reads every row of the vocabulary table
        Wm = arena.s(wname)
        n = x.size(0)
        N, K = (Wm.size(1), Wm.size(0)) if transposed else (Wm.size(0), Wm.size(1))
        x = x.contiguous()
        bias = arena.m(bname) if bname else None
        pre = None
        if out_f32:
            # row pitch padded to 8 elements so the bf16 copy of d(logits) is a legal TMA operand
            out = torch.empty((n, _pad8(N)), dtype=F32, device=x.device)[:, :N]
            if n:
                _lib.gemm(x, Wm, n, N, K, b_mn=transposed, bias=bias, act=act, out_f32=out)
        else:
            out = torch.empty((n, N), dtype=BF16, device=x.device)
            if act == _lib.ACT_GELU:
                pre = torch.empty((n, N), dtype=BF16, device=x.device)
            if n:
                _lib.gemm(x, Wm, n, N, K, b_mn=transposed, bias=bias, act=act, out_bf16=out, out_pre=pre)
        ctx.save_for_backward(x, pre, out if act == _lib.ACT_TANH else None)
        ctx.meta = (arena, wname, bname, act, transposed, N, K)
        return out

    @staticmethod
    def backward(ctx, dy):
        arena, wname, bname, act, transposed, N, K = ctx.meta
        x, pre, _ = ctx.saved_tensors
        n = x.size(0)
        if n == 0:
            return torch.zeros_like(x), None, None, None, None, None, None
        dy = _as_bf16_rows(dy)
        if act == _lib.ACT_GELU:
            dz = _gelu_bwd(dy, pre)                   # dz = dy * gelu'(pre)
        elif act == _lib.ACT_NONE:
            dz = dy
        else:
            raise RuntimeError("LinearFn.backward: unsupported activation")
        return _linear_backward(arena, wname, bname, x, dz, transposed, N, K), None, None, None, None, None, None


def _linear_backward(arena, wname, bname, x, dz, transposed, N, K):
    """dx of y = x W^T (+ b) given dz = dL/dy (bf16, 8-element pitch); accumulates dW and db into the arena."""
    n = x.size(0)
    Wm = arena.s(wname)
    if bname:
        call("uc2_colsum_bf16", dz.data_ptr(), dz.stride(0), n, N, arena.gp(bname), stream())
    big = N > 8192       # vocabulary-sized contraction: split it over the SMs, accumulate in fp32
    dx32 = torch.zeros((n, K), dtype=F32, device=x.device) if big else None
    dx = torch.empty((n, K), dtype=BF16, device=x.device)
    okw = dict(out_f32=dx32, accumulate=True, split_k=0) if big else dict(out_bf16=dx)
    if not transposed:
        _lib.gemm(dz, Wm, n, K, N, b_mn=True, **okw)                             # dx = dz W
        _lib.gemm(dz, x, N, K, n, a_mn=True, b_mn=True, out_f32=arena.g(wname), accumulate=True, split_k=0)
    else:
        _lib.gemm(dz, Wm, n, K, N, **okw)                                        # dx = dz (W^T)^T, W is [K,N]
        _lib.gemm(x, dz, K, N, n, a_mn=True, b_mn=True, out_f32=arena.g(wname), accumulate=True, split_k=0)
    if big:
        call("uc2_cast_f32_bf16", dx32.data_ptr(), dx.data_ptr(), dx.numel(), stream())
    if wname.endswith("embeddings.word_embeddings.weight"):
        arena.word_emb_dense = True          # tied decoder: the vocabulary-table gradient is dense this step
        sync = getattr(arena, "grad_sync", None)
        if (sync is not None and sync.enabled and getattr(sync, "early_dense", False) and arena.offset[wname] == 0
                and not getattr(arena, "word_dense_sent", False)):
            # data parallel: this dense term is complete now, the lookup's rows come at the very end of the backward
            # pass -> average it under the encoder's backward; embed backward sends its rows separately (side_rows)
            lo = arena.offset[wname]
            sync.ready(lo, lo + arena.numel[wname])
            arena.word_dense_sent = True
    arena.touch(wname, *( [bname] if bname else []))
    return dx


class LmHeadCEFn(torch.autograd.Function):
    """Tied MLM decoder + cross entropy as one autograd node (model/layer.py:263-264 + model/model.py:592-596).
    Forward: ONE GEMM writes the logits h W_emb^T + bias as bf16 and, from the fp32 accumulators in its epilogue, a
    (max, sum-exp) pair per row and 32-column chunk plus the label's logit (csrc/gemm_tcgen05.cu MODE 9);
    uc2_ce_stats_reduce merges them into lse and the per-row loss.  No fp32 [n, 250 002] tensor exists.  Backward
    turns the bf16 logits into d(logits) in place (uc2_ce_bwd_inplace_bf16), which is already the GEMM operand of the
    decoder's dgrad / wgrad."""

    @staticmethod
    def forward(ctx, h, arena, wname, bname, targets, ignore_index):
        lazy = getattr(arena, "lazy", None)
        if lazy is not None and wname == lazy.name:
            lazy.catch_up_all()             # the tied decoder reads every row of the table
        Wm = arena.s(wname)
        n, K = h.shape
        N = Wm.size(0)
        h = h.contiguous()
        dev = h.device
        logits = torch.empty((n, _pad16(N)), dtype=BF16, device=dev)
        loss = torch.empty((n,), dtype=F32, device=dev)
        lse = torch.empty((n,), dtype=F32, device=dev)
        t = targets.to(torch.long).contiguous()
        if n:
            n_chunks = (N + 31) // 32
            stats = torch.empty((n_chunks, n, 2), dtype=F32, device=dev)
            part = torch.empty(((n_chunks + 255) // 256, n, 2), dtype=F32, device=dev)
            tgt = torch.zeros((n,), dtype=F32, device=dev)       # rows with an out-of-range (ignored) label keep 0
            _lib.gemm(h, Wm, n, N, K, bias=arena.m(bname), out_bf16=logits[:, :N], ld_out=logits.stride(0),
                      ce=(stats, t, tgt))
            call("uc2_ce_stats_reduce", stats.data_ptr(), n, n_chunks, n, tgt.data_ptr(), t.data_ptr(), ignore_index,
                 part.data_ptr(), loss.data_ptr(), lse.data_ptr(), stream())
        ctx.save_for_backward(h, logits, lse, t)
        ctx.meta = (arena, wname, bname, N, K, ignore_index)
        return loss

    @staticmethod
    def backward(ctx, dloss):
        arena, wname, bname, N, K, ignore_index = ctx.meta
        h, logits, lse, t = ctx.saved_tensors
        n = h.size(0)
        if n == 0:
            return torch.zeros_like(h), None, None, None, None, None
        dloss = dloss.to(F32).contiguous()
        # in place: the saved logits become d(logits) (this node's backward runs once)
        call("uc2_ce_bwd_inplace_bf16", logits.data_ptr(), logits.stride(0), n, N, t.data_ptr(), ignore_index,
             dloss.data_ptr(), lse.data_ptr(), stream())
        return _linear_backward(arena, wname, bname, h, logits[:, :N], False, N, K), None, None, None, None, None


def _pad16(n):
    return (n + 15) // 16 * 16


def _pad8(n):
    return (n + 7) // 8 * 8


def _as_bf16_rows(t):
    """[n, N] gradient -> bf16 with unit inner stride and a row pitch that is a multiple of 8 elements."""
    n, N = t.shape
    if t.dtype == BF16 and t.stride(1) == 1 and t.stride(0) % 8 == 0 and t.data_ptr() % 16 == 0:
        return t
    out = torch.empty((n, _pad8(N)), dtype=BF16, device=t.device)[:, :N]
    if n == 0:
        return out
    if t.dtype == F32 and t.stride(1) == 1:
        call("uc2_f32_to_bf16_2d", t.data_ptr(), t.stride(0), out.data_ptr(), out.stride(0), n, N, stream())
    else:
        out.copy_(t)
    return out


def _gelu_bwd(dy, pre):
    """dz = dy * gelu'(pre) via the GEMM-free elementwise kernel (head rows only, small)."""
    out = torch.empty_like(dy)
    call("uc2_dgelu_bf16", dy.data_ptr(), pre.data_ptr(), out.data_ptr(), dy.numel(), stream())
    return out


class LayerNormFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, arena, wname, bname, eps):
        x = x.contiguous()
        y = torch.empty_like(x)
        if x.size(0):
            call("uc2_layernorm_fwd", x.data_ptr(), 0, arena.mp(wname), arena.mp(bname), eps, y.data_ptr(), None,
                 x.size(0), stream())
        ctx.save_for_backward(x)
        ctx.meta = (arena, wname, bname, eps)
        return y

    @staticmethod
    def backward(ctx, dy):
        arena, wname, bname, eps = ctx.meta
        (x,) = ctx.saved_tensors
        dy = dy.to(BF16).contiguous()
        dx = torch.empty_like(x)
        if x.size(0):
            call("uc2_layernorm_bwd", x.data_ptr(), 0, dy.data_ptr(), arena.mp(wname), eps, dx.data_ptr(),
                 arena.gp(wname), arena.gp(bname), None, x.size(0), stream())
        arena.touch(wname, bname)
        return dx, None, None, None, None


_PENDING_COUNTS = []        # (device count, host hint) pairs of MaskedRowsFn calls not yet compared


def verify_masked_counts():
    """The reference's boolean indexing cannot disagree with its own mask; MaskedRowsFn takes the number of set entries as
    a HOST hint (so the step has no device-to-host sync) and the device scan counts them again.  This compares every
    pair recorded since the last call (one sync) and raises on a mismatch -- a wrong hint would silently drop rows or
    append zero rows.  The training loops call it whenever they read their logged losses back."""
    global _PENDING_COUNTS
    pending, _PENDING_COUNTS = _PENDING_COUNTS, []
    if pending:
        got = torch.stack([c.reshape(()) for c, _ in pending]).cpu().tolist()
        for g, (_, hint) in zip(got, pending):
            if int(g) != int(hint):
                raise RuntimeError(f"masked-row compaction: the batch says {hint} masked positions (n_masked / target rows) "
                                   f"but the mask holds {g}")


class MaskedRowsFn(torch.autograd.Function):
    """hidden[:, :L][mask] in row-major order (model.py:653-657).  `count` is the number of set mask entries
    (known on the host from the batch, or computed with one sync like the reference's boolean indexing); the device
    count is checked against it later (verify_masked_counts), and rows beyond the device count stay zero."""

    @staticmethod
    def forward(ctx, hidden, mask, count):
        B, S, H = hidden.shape
        L = mask.size(1)
        m8 = mask.to(torch.uint8).contiguous()
        idx = torch.empty(max(count, 1), dtype=torch.int32, device=hidden.device)
        cnt = torch.empty(1, dtype=torch.int32, device=hidden.device)
        call("uc2_mask_scan", m8.data_ptr(), m8.numel(), idx.data_ptr(), cnt.data_ptr(), max(count, 1), stream())
        out = torch.zeros((count, H), dtype=BF16, device=hidden.device)
        hidden = hidden.contiguous()
        if count:
            call("uc2_gather_rows", hidden.data_ptr(), idx.data_ptr(), cnt.data_ptr(), L, S, out.data_ptr(), count, stream())
        _PENDING_COUNTS.append((cnt, count))
        if len(_PENDING_COUNTS) >= 256:
            verify_masked_counts()
        ctx.save_for_backward(idx, cnt)
        ctx.meta = (B, S, H, L, count)
        return out

    @staticmethod
    def backward(ctx, dout):
        idx, cnt = ctx.saved_tensors
        B, S, H, L, count = ctx.meta
        dh = torch.zeros((B, S, H), dtype=BF16, device=dout.device)
        if count:
            dout = dout.to(BF16).contiguous()
            call("uc2_scatter_rows_add", dout.data_ptr(), idx.data_ptr(), cnt.data_ptr(), L, S, dh.data_ptr(), count, stream())
        return dh, None, None


class PoolerFn(torch.autograd.Function):
    """BertPooler (model/layer.py:179-185): tanh(dense(hidden[:, 0])) -> fp32 [B,768]."""

    @staticmethod
    def forward(ctx, hidden, arena, wname, bname):
        B, S, H = hidden.shape
        hidden = hidden.contiguous()
        out = torch.empty((B, H), dtype=F32, device=hidden.device)
        _lib.gemm(hidden, arena.s(wname), B, H, H, lda=S * H, bias=arena.m(bname), act=_lib.ACT_TANH, out_f32=out)
        ctx.save_for_backward(hidden, out)
        ctx.meta = (arena, wname, bname)
        return out

    @staticmethod
    def backward(ctx, dy):
        arena, wname, bname = ctx.meta
        hidden, y = ctx.saved_tensors
        B, S, H = hidden.shape
        dy = dy.to(F32).contiguous()
        dpre = torch.empty((B, H), dtype=BF16, device=dy.device)
        call("uc2_tanh_bwd", y.data_ptr(), dy.data_ptr(), dpre.data_ptr(), B * H, stream())
        call("uc2_colsum_bf16", dpre.data_ptr(), H, B, H, arena.gp(bname), stream())
        # dW[H,H] += dpre^T h0   (h0 = hidden[:,0], row pitch S*H)
        _lib.gemm(dpre, hidden, H, H, B, a_mn=True, b_mn=True, ldb=S * H, out_f32=arena.g(wname), accumulate=True,
                  split_k=0)
        dh = torch.zeros((B, S, H), dtype=BF16, device=dy.device)
        _lib.gemm(dpre, arena.s(wname), B, H, H, b_mn=True, out_bf16=dh, ld_out=S * H)
        arena.touch(wname, bname)
        return dh, None, None, None


class NarrowLinearFn(torch.autograd.Function):
    """Linear(768, N<=8) on fp32 rows: itm_output / rank_output."""

    @staticmethod
    def forward(ctx, x, arena, wname, bname):
        x = x.contiguous()
        M, K = x.shape
        N = arena.shape[wname][0]
        out = torch.empty((M, N), dtype=F32, device=x.device)
        call("uc2_narrow_linear_fwd", x.data_ptr(), K, arena.mp(wname), arena.mp(bname), out.data_ptr(), M, N, K, stream())
        ctx.save_for_backward(x)
        ctx.meta = (arena, wname, bname, N)
        return out

    @staticmethod
    def backward(ctx, dy):
        arena, wname, bname, N = ctx.meta
        (x,) = ctx.saved_tensors
        M, K = x.shape
        dy = dy.to(F32).contiguous()
        dx = torch.empty_like(x)
        call("uc2_narrow_linear_bwd", x.data_ptr(), K, arena.mp(wname), dy.data_ptr(), dx.data_ptr(), K, arena.gp(wname),
             arena.gp(bname), M, N, K, stream())
        arena.touch(wname, bname)
        return dx, None, None, None


class RankLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, scores, sample_size, margin):
        s = scores.contiguous().view(-1)
        groups = s.numel() // sample_size
        loss = torch.empty((groups, sample_size - 1), dtype=F32, device=s.device)
        call("uc2_rank_loss_fwd", s.data_ptr(), loss.data_ptr(), groups, sample_size, margin, stream())
        ctx.save_for_backward(s)
        ctx.meta = (groups, sample_size, margin, scores.shape)
        return loss

    @staticmethod
    def backward(ctx, dloss):
        (s,) = ctx.saved_tensors
        groups, ss, margin, shape = ctx.meta
        dloss = dloss.to(F32).contiguous()
        ds = torch.empty_like(s)
        call("uc2_rank_loss_bwd", s.data_ptr(), dloss.data_ptr(), ds.data_ptr(), groups, ss, margin, stream())
        return ds.view(shape), None, None


class SoftmaxLossFn(torch.autograd.Function):
    """kind 0: cross entropy (targets int64, ignore_index); kind 1: KL vs soft targets (elementwise)."""

    @staticmethod
    def forward(ctx, logits, kind, targets, ignore_index):
        if logits.stride(1) != 1:
            logits = logits.contiguous()
        n, Cc = logits.shape
        ld = logits.stride(0) if n > 1 else max(Cc, logits.stride(0))
        if kind == 0:
            t = targets.to(torch.long).contiguous()
            loss = torch.empty((n,), dtype=F32, device=logits.device)
            call("uc2_softmax_loss", logits.data_ptr(), ld, n, Cc, 0, t.data_ptr(), ignore_index, None, loss.data_ptr(),
                 None, None, None, stream())
        else:
            t = targets.to(F32).contiguous()
            loss = torch.empty((n, Cc), dtype=F32, device=logits.device)
            call("uc2_softmax_loss", logits.data_ptr(), ld, n, Cc, 1, None, -1, t.data_ptr(), loss.data_ptr(), None, None,
                 None, stream())
        ctx.save_for_backward(logits, t)
        ctx.meta = (kind, ignore_index, ld)
        return loss

    @staticmethod
    def backward(ctx, dloss):
        logits, t = ctx.saved_tensors
        kind, ignore_index, ld = ctx.meta
        n, Cc = logits.shape
        dloss = dloss.to(F32).contiguous()
        dl = torch.empty((n, ld), dtype=F32, device=logits.device)[:, :Cc]     # same row pitch as the logits
        if kind == 0:
            call("uc2_softmax_loss", logits.data_ptr(), ld, n, Cc, 0, t.data_ptr(), ignore_index, None, None,
                 dloss.data_ptr(), dl.data_ptr(), None, stream())
        else:
            call("uc2_softmax_loss", logits.data_ptr(), ld, n, Cc, 1, None, -1, t.data_ptr(), None, dloss.data_ptr(),
                 dl.data_ptr(), None, stream())
        return dl, None, None, None


class SelectColumnsFn(torch.autograd.Function):
    """logits[:, cols] (model/model.py:643 `prediction[:, VALID_XLMR_TOKEN_IDS]`): pure index plumbing -- a column
    gather forward, a column scatter backward; the output row pitch is padded like the logits'."""

    @staticmethod
    def forward(ctx, logits, cols):
        n = logits.size(0)
        out = torch.empty((n, _pad8(cols.numel())), dtype=logits.dtype, device=logits.device)[:, :cols.numel()]
        torch.index_select(logits, 1, cols, out=out) if out.is_contiguous() else out.copy_(logits.index_select(1, cols))
        ctx.save_for_backward(cols)
        ctx.shape = (n, logits.size(1), logits.stride(0))
        return out

    @staticmethod
    def backward(ctx, dout):
        (cols,) = ctx.saved_tensors
        n, Cc, ld = ctx.shape
        d = torch.zeros((n, ld), dtype=dout.dtype, device=dout.device)[:, :Cc]
        d.index_add_(1, cols, dout)
        return d, None


class MseFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, target):
        pred, target = pred.contiguous(), target.to(F32).contiguous()
        loss = torch.empty_like(pred)
        call("uc2_mse", pred.data_ptr(), target.data_ptr(), loss.data_ptr(), None, None, pred.numel(), stream())
        ctx.save_for_backward(pred, target)
        return loss

    @staticmethod
    def backward(ctx, dloss):
        pred, target = ctx.saved_tensors
        dloss = dloss.to(F32).contiguous()
        dp = torch.empty_like(pred)
        call("uc2_mse", pred.data_ptr(), target.data_ptr(), None, dloss.data_ptr(), dp.data_ptr(), pred.numel(), stream())
        return dp, None
