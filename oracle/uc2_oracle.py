"""ORACLE -- test infrastructure, NOT product code.

A CPU fp32 restatement of the UC2 cross-modal encoder hot path, written as pure
functions over a ``{state_dict name: tensor}`` mapping.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs may import this file, and only as the checker or the timed CPU baseline.
The product path (``uc2_b200``) never imports it.

Pinning: the reference repository holds no golden vectors or tests for this
path.  The oracle is pinned instead against outputs
of the reference's own modules executed in the authoring container
(``tests/golden/make_golden.py`` imports /root/reference with the shims SURVEY 8c
lists and stores their outputs under ``tests/golden/*.npz``;
``tests/test_oracle_golden.py`` replays them).  apex FusedLayerNorm is shimmed to
``torch.nn.LayerNorm`` per BASELINE.json configs[0]; parity at that one boundary
(and for apex-amp / Horovod semantics) is therefore "unpinned", see DESIGN.md.

Integer/index functions are numpy; floating-point functions are torch fp32 on
CPU (autograd supplies the reference gradients).  Every function cites the
reference file:line (relative to /root/reference) it follows.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

H = 768


# ----------------------------------------------------------------------------
# integer / index work (bit-exact contracts)
# ----------------------------------------------------------------------------
def position_ids_from_input_ids(input_ids, padding_idx=1):
    """model/model.py:280-290 -- cumsum over non-pad tokens, pads keep padding_idx."""
    ids = np.asarray(input_ids)
    out = np.empty_like(ids, dtype=np.int64)
    for b in range(ids.shape[0]):
        run = 0
        for t in range(ids.shape[1]):
            if ids[b, t] != padding_idx:
                run += 1
                out[b, t] = run + padding_idx
            else:
                out[b, t] = padding_idx
    return out


def gather_index(txt_lens, num_bbs, max_len, out_size):
    """data/data.py:376-384."""
    gi = np.tile(np.arange(out_size, dtype=np.int64), (len(txt_lens), 1))
    for i, (tl, nbb) in enumerate(zip(txt_lens, num_bbs)):
        for k in range(nbb):
            gi[i, tl + k] = max_len + k
    return gi


def ot_scatter(txt_lens, max_txt_len, joint_len):
    """data/itm.py:264-271."""
    sc = np.tile(np.arange(joint_len, dtype=np.int64), (len(txt_lens), 1))
    for i, tl in enumerate(txt_lens):
        for k in range(joint_len - tl):
            sc[i, tl + k] = max_txt_len + k
    return sc


def pad_mask(lens, max_len):
    """data/itm.py:274-278 (as bool)."""
    p = np.zeros((len(lens), max_len), dtype=bool)
    for i, l in enumerate(lens):
        p[i, l:] = True
    return p


def masked_row_order(mask):
    """Row-major (b, then j) positions selected by hidden[mask] (model/model.py:653-657)."""
    m = np.asarray(mask).astype(bool)
    return [(b, j) for b in range(m.shape[0]) for j in range(m.shape[1]) if m[b, j]]


# ----------------------------------------------------------------------------
# floating-point building blocks
# ----------------------------------------------------------------------------
def gelu_erf(x):
    """model/layer.py:31-37."""
    return x * 0.5 * (1.0 + torch.erf(x / math.sqrt(2.0)))


def layer_norm(x, w, b, eps):
    """apex FusedLayerNorm shimmed to torch LayerNorm (BASELINE.json configs[0])."""
    return F.layer_norm(x, (x.shape[-1],), w, b, eps)


class Family:
    """Switches between the VLXLMR (default, what pretrain.py/itm.py run) and Uniter
    families -- SURVEY 8a row A0."""

    def __init__(self, name="vlxlmr", layer_norm_eps=1e-5):
        self.name = name
        if name == "vlxlmr":
            self.enc = "roberta."
            self.type_emb = "embeddings.new_token_type_embeddings.weight"
            self.emb_eps = layer_norm_eps
            self.pad_id = 1
        else:
            self.enc = "bert."
            self.type_emb = "embeddings.token_type_embeddings.weight"
            self.emb_eps = 1e-12
            self.pad_id = 0


def text_embeddings(sd, fam, input_ids, position_ids=None):
    """model/model.py:304-335 (VLXLMR) / 987-1001 (Uniter); dropout omitted (p=0)."""
    p = fam.enc
    ids = torch.as_tensor(input_ids)
    if position_ids is None:
        assert fam.name == "vlxlmr"
        position_ids = torch.from_numpy(position_ids_from_input_ids(ids.numpy(), fam.pad_id))
    position_ids = torch.as_tensor(position_ids)
    if position_ids.dim() == 2 and position_ids.size(0) == 1:
        position_ids = position_ids.expand(ids.size(0), -1)
    e = (sd[p + "embeddings.word_embeddings.weight"][ids]
         + sd[p + "embeddings.position_embeddings.weight"][position_ids]
         + sd[p + fam.type_emb][0])
    return layer_norm(e, sd[p + "embeddings.LayerNorm.weight"], sd[p + "embeddings.LayerNorm.bias"],
                      fam.emb_eps)


def image_embeddings(sd, fam, img_feat, img_pos_feat, img_masks=None):
    """model/model.py:352-364 + caller 401-410 (type id 1 for every region)."""
    p = fam.enc + "img_embeddings."
    if img_masks is not None:
        me = sd[p + "mask_embedding.weight"]
        # row 0 is forced to zero on every call (model.py:354); masked regions add row 1
        img_feat = img_feat + torch.as_tensor(img_masks).bool().unsqueeze(-1).to(img_feat.dtype) * me[1]
    t_im = layer_norm(F.linear(img_feat, sd[p + "img_linear.weight"], sd[p + "img_linear.bias"]),
                      sd[p + "img_layer_norm.weight"], sd[p + "img_layer_norm.bias"], fam.emb_eps)
    t_pos = layer_norm(F.linear(img_pos_feat, sd[p + "pos_linear.weight"], sd[p + "pos_linear.bias"]),
                       sd[p + "pos_layer_norm.weight"], sd[p + "pos_layer_norm.bias"], fam.emb_eps)
    e = t_im + t_pos + sd[fam.enc + fam.type_emb][1]
    return layer_norm(e, sd[p + "LayerNorm.weight"], sd[p + "LayerNorm.bias"], fam.emb_eps)


def pack(txt_emb, img_emb, gi):
    """model/model.py:421-424: out[b,j] = cat(txt,img)[b, gather_index[b,j]]."""
    cat = torch.cat([txt_emb, img_emb], 1)
    gi = torch.as_tensor(gi)
    return torch.gather(cat, 1, gi.unsqueeze(-1).expand(-1, -1, cat.size(-1)))


def self_attention(sd, pre, x, ext_mask, n_heads=12, drop_attn=None):
    """model/layer.py:75-101.  drop_attn: optional [B, heads, S, S] multiplier (0 or 1/(1-p)) standing in for
    nn.Dropout on the attention probabilities (layer.py:94)."""
    B, S, Hd = x.shape
    d = Hd // n_heads

    def proj(n):
        y = F.linear(x, sd[pre + f"attention.self.{n}.weight"], sd[pre + f"attention.self.{n}.bias"])
        return y.view(B, S, n_heads, d).permute(0, 2, 1, 3)
    q, k, v = proj("query"), proj("key"), proj("value")
    s = q @ k.transpose(-1, -2) / math.sqrt(d) + ext_mask
    pr = torch.softmax(s, -1)
    if drop_attn is not None:
        pr = pr * drop_attn
    return (pr @ v).permute(0, 2, 1, 3).reshape(B, S, Hd)


def bert_layer(sd, pre, x, ext_mask, n_heads=12, drop=None):
    """model/layer.py:159-170 (A6->A7->A8->A9); LayerNorm eps is 1e-12 in both families
    (layer.py:108,149).  drop: optional (attn [B,h,S,S], out1 [B,S,H], out2 [B,S,H]) multipliers for the three
    nn.Dropout sites (layer.py:94, 113, 154); None = evaluation / p = 0."""
    d_attn, d_out1, d_out2 = drop if drop is not None else (None, None, None)
    ctx = self_attention(sd, pre, x, ext_mask, n_heads, d_attn)
    o1 = F.linear(ctx, sd[pre + "attention.output.dense.weight"], sd[pre + "attention.output.dense.bias"])
    if d_out1 is not None:
        o1 = o1 * d_out1
    a = layer_norm(o1 + x, sd[pre + "attention.output.LayerNorm.weight"],
                   sd[pre + "attention.output.LayerNorm.bias"], 1e-12)
    i = gelu_erf(F.linear(a, sd[pre + "intermediate.dense.weight"], sd[pre + "intermediate.dense.bias"]))
    o2 = F.linear(i, sd[pre + "output.dense.weight"], sd[pre + "output.dense.bias"])
    if d_out2 is not None:
        o2 = o2 * d_out2
    return layer_norm(o2 + a, sd[pre + "output.LayerNorm.weight"], sd[pre + "output.LayerNorm.bias"], 1e-12)


def n_layers(sd, fam):
    n = 0
    while (fam.enc + f"encoder.layer.{n}.output.dense.weight") in sd:
        n += 1
    return n


def encoder(sd, fam, input_ids, position_ids, img_feat, img_pos_feat, attention_mask,
            gather_index=None, img_masks=None, all_layers=False, drop=None):
    """model/model.py:427-458 == 1109-1140.  drop: optional dict of dropout multipliers (training mode with the
    masks given explicitly): "emb" [B, T+R, H] on cat(text, image) embeddings before the pack (model.py:334, 363),
    "layers": list of (attn, out1, out2) per layer."""
    am = torch.as_tensor(attention_mask)
    ext = (1.0 - am[:, None, None, :].to(torch.float32)) * -10000.0
    if input_ids is None:
        x = image_embeddings(sd, fam, img_feat, img_pos_feat, img_masks)
    elif img_feat is None:
        x = text_embeddings(sd, fam, input_ids, position_ids)
    else:
        te = text_embeddings(sd, fam, input_ids, position_ids)
        ie = image_embeddings(sd, fam, img_feat, img_pos_feat, img_masks)
        if drop is not None and drop.get("emb") is not None:
            T = te.size(1)
            te, ie = te * drop["emb"][:, :T], ie * drop["emb"][:, T:]
        x = pack(te, ie, gather_index)
    outs = [x]
    for l in range(n_layers(sd, fam)):
        x = bert_layer(sd, fam.enc + f"encoder.layer.{l}.", x, ext,
                       drop=drop["layers"][l] if drop is not None else None)
        outs.append(x)
    return outs if all_layers else x


def pooler(sd, fam, h):
    """model/layer.py:179-185."""
    return torch.tanh(F.linear(h[:, 0], sd[fam.enc + "pooler.dense.weight"], sd[fam.enc + "pooler.dense.bias"]))


def masked_hidden(hidden, mask):
    """model/model.py:653-657."""
    return hidden[torch.as_tensor(mask).bool()]


# ----------------------------------------------------------------------------
# heads
# ----------------------------------------------------------------------------
def mlm_head(sd, fam, x):
    """RobertaLMHead layer.py:257-265 (decoder tied to word embeddings, bias = cls.bias)
    / BertLMPredictionHead layer.py:199-222."""
    W = sd[fam.enc + "embeddings.word_embeddings.weight"]
    if fam.name == "vlxlmr":
        h = layer_norm(gelu_erf(F.linear(x, sd["cls.dense.weight"], sd["cls.dense.bias"])),
                       sd["cls.layer_norm.weight"], sd["cls.layer_norm.bias"], fam.emb_eps)
        return F.linear(h, W, sd["cls.bias"])
    t = "cls.predictions.transform."
    h = layer_norm(gelu_erf(F.linear(x, sd[t + "dense.weight"], sd[t + "dense.bias"])),
                   sd[t + "LayerNorm.weight"], sd[t + "LayerNorm.bias"], 1e-12)
    return F.linear(h, W) + sd["cls.predictions.bias"]


def mrfr_head(sd, fam, x):
    """RegionFeatureRegression model/model.py:1143-1156 (output weight tied to img_linear)."""
    h = layer_norm(gelu_erf(F.linear(x, sd["feat_regress.net.0.weight"], sd["feat_regress.net.0.bias"])),
                   sd["feat_regress.net.2.weight"], sd["feat_regress.net.2.bias"], 1e-12)
    return F.linear(h, sd[fam.enc + "img_embeddings.img_linear.weight"].t(), sd["feat_regress.bias"])


def mrc_head(sd, x):
    """RegionClassification model/model.py:1159-1169."""
    h = layer_norm(gelu_erf(F.linear(x, sd["region_classifier.net.0.weight"],
                                     sd["region_classifier.net.0.bias"])),
                   sd["region_classifier.net.2.weight"], sd["region_classifier.net.2.bias"], 1e-12)
    return F.linear(h, sd["region_classifier.net.3.weight"], sd["region_classifier.net.3.bias"])


# ----------------------------------------------------------------------------
# optimal transport (model/ot.py)
# ----------------------------------------------------------------------------
def cost_matrix_cosine(x, y, eps=1e-5):
    """ot.py:8-18."""
    xn = x / x.norm(dim=-1, keepdim=True).clamp_min(eps)
    yn = y / y.norm(dim=-1, keepdim=True).clamp_min(eps)
    return 1 - xn @ yn.transpose(1, 2)


@torch.no_grad()
def ipot(C, x_len, x_pad, y_len, y_pad, joint_pad, beta=0.5, iteration=50, k=1):
    """ot.py:32-63. C [B,M,N]; returns T [B,N,M]."""
    b, m, n = C.shape
    sigma = torch.ones(b, m, dtype=C.dtype) / x_len[:, None]
    T = torch.ones(b, n, m, dtype=C.dtype)
    A = torch.exp(-C.transpose(1, 2) / beta)
    jp = joint_pad.transpose(1, 2)
    sigma = sigma.masked_fill(x_pad, 0)
    T = T.masked_fill(jp, 0)
    A = A.masked_fill(jp, 0)
    xl, yl = x_len[:, None, None], y_len[:, None, None]
    x_mask = (x_pad.to(C.dtype) * 1e4)[:, None, :]
    y_mask = (y_pad.to(C.dtype) * 1e4)[:, None, :]
    for _ in range(iteration):
        Q = A * T
        sigma = sigma.view(b, m, 1)
        for _ in range(k):
            delta = 1 / (yl * (Q @ sigma).view(b, 1, n) + y_mask)
            sigma = 1 / (xl * (delta @ Q) + x_mask)
        T = delta.view(b, n, 1) * Q * sigma
    return T.masked_fill(jp, 0)


def optimal_transport_dist(txt_emb, img_emb, txt_pad, img_pad, beta=0.5, iteration=50, k=1):
    """ot.py:66-82; trace(C @ T) == sum_{m,n} C[m,n] T[n,m] (ot.py:21-29)."""
    txt_pad, img_pad = torch.as_tensor(txt_pad).bool(), torch.as_tensor(img_pad).bool()
    cost = cost_matrix_cosine(txt_emb, img_emb)
    joint_pad = txt_pad[:, :, None] | img_pad[:, None, :]
    cost = cost.masked_fill(joint_pad, 0)
    txt_len = (txt_pad.size(1) - txt_pad.sum(1)).to(cost.dtype)
    img_len = (img_pad.size(1) - img_pad.sum(1)).to(cost.dtype)
    T = ipot(cost.detach(), txt_len, txt_pad, img_len, img_pad, joint_pad, beta, iteration, k)
    return (cost @ T.detach()).diagonal(dim1=-2, dim2=-1).sum(-1)


# ----------------------------------------------------------------------------
# task forwards (model/model.py:495-775, model/itm.py:28-55)
# ----------------------------------------------------------------------------
def forward_pretraining(sd, fam, batch, task, compute_loss=True, ot_pos_only=False, drop=None):
    """drop: optional dropout multipliers for the encoder (see encoder()); the heads have no dropout."""
    ids = batch["input_ids"]
    pos = batch.get("position_ids") if (task == "tlm" or fam.name != "vlxlmr") else None
    feat, posf = batch.get("img_feat"), batch.get("img_pos_feat")
    am, gi = batch["attn_masks"], batch.get("gather_index")
    if task in ("mlm", "tlm", "tlm-ni"):
        if task == "tlm-ni":
            feat = posf = gi = None
        h = encoder(sd, fam, ids, pos, feat, posf, am, gi, drop=drop)
        h = h[:, :ids.size(1)]                                   # model.py:583
        lab = batch["txt_labels"]
        scores = mlm_head(sd, fam, masked_hidden(h, lab != -1))
        return F.cross_entropy(scores, lab[lab != -1], reduction="none") if compute_loss else scores
    if task in ("mmxlm", "vmlm"):                                 # model.py:598-624
        h = encoder(sd, fam, ids, pos, feat, posf, am, gi, batch["img_masks"], drop=drop)
        lab = batch["txt_labels"]
        scores = mlm_head(sd, fam, masked_hidden(h, lab != -1))
        return F.cross_entropy(scores, lab[lab != -1], reduction="none") if compute_loss else scores
    if task in ("mmxlm-soft", "vmlm-soft"):                       # model.py:626-651
        h = encoder(sd, fam, ids, pos, feat, posf, am, gi, batch["img_masks"], drop=drop)
        pred = mlm_head(sd, fam, masked_hidden(h, batch["tgt_masks"]))[:, torch.as_tensor(batch["valid_token_ids"])]
        if not compute_loss:
            return pred
        return F.kl_div(F.log_softmax(pred, -1), batch["label_targets"], reduction="none")
    if task == "mrfr":
        h = encoder(sd, fam, ids, pos, feat, posf, am, gi, batch["img_masks"], drop=drop)
        pred = mrfr_head(sd, fam, masked_hidden(h, batch["img_mask_tgt"]))
        return F.mse_loss(pred, batch["feat_targets"], reduction="none") if compute_loss else pred
    if task.startswith("mrc"):
        h = encoder(sd, fam, ids, pos, feat, posf, am, gi, batch["img_masks"], drop=drop)
        pred = mrc_head(sd, masked_hidden(h, batch["img_mask_tgt"]))
        if not compute_loss:
            return pred
        if "kl" in task:
            return F.kl_div(F.log_softmax(pred, -1), batch["label_targets"], reduction="none")
        tgt = batch["label_targets"][:, 1:].max(-1)[1] + 1
        return F.cross_entropy(pred, tgt, ignore_index=0, reduction="none")
    if task == "itm":
        h = encoder(sd, fam, ids, pos, feat, posf, am, gi, drop=drop)
        scores = F.linear(pooler(sd, fam, h), sd["itm_output.weight"], sd["itm_output.bias"])
        targets = batch["targets"]
        ot = batch.get("ot_inputs")
        ot_loss = None
        if ot is not None:
            tl, il = ids.size(1), feat.size(1)
            max_l = max(ot["scatter_max"] + 1, tl + il)
            sc = ot["ot_scatter"].unsqueeze(-1).expand_as(h)
            ctx = torch.zeros(h.size(0), max_l, h.size(-1), dtype=h.dtype).scatter(1, sc, h)
            dist = optimal_transport_dist(ctx[:, :tl], ctx[:, tl:tl + il], ot["txt_pad"], ot["img_pad"])
            ot_loss = dist[targets == 1] if ot_pos_only else (dist[targets == 1], dist[targets == 0])
        if compute_loss:
            return F.cross_entropy(scores, targets, reduction="none"), ot_loss
        return scores, ot_loss
    raise ValueError("invalid task")


def pretraining_loss(out, task, itm_ot_lambda=0.1, ot_pos_only=False):
    """Driver-side reduction, pretrain.py:524-553."""
    if task.startswith("itm"):
        itm, ot = out
        loss = itm.mean()
        if ot is not None:
            if ot_pos_only:
                loss = loss + itm_ot_lambda * ot.mean()
            else:
                p, n = ot
                loss = loss + itm_ot_lambda * (p.sum() - n.sum()) / (p.size(0) + n.size(0))
        return loss
    return out.mean()


def forward_retrieval(sd, fam, batch, compute_loss=True, margin=0.2, drop=None):
    """model/itm.py:28-55."""
    pos = None if fam.name == "vlxlmr" else batch.get("position_ids")
    h = encoder(sd, fam, batch["input_ids"], pos, batch["img_feat"], batch["img_pos_feat"],
                batch["attn_masks"], batch["gather_index"], drop=drop)
    scores = F.linear(pooler(sd, fam, h), sd["rank_output.weight"], sd["rank_output.bias"])
    if not compute_loss:
        return scores
    s = torch.sigmoid(scores).contiguous().view(-1, batch["sample_size"])
    return torch.clamp(margin + s[:, 1:] - s[:, :1], 0)


# ----------------------------------------------------------------------------
# optimiser (optim/adamw.py:40-103, optim/misc.py:9-32, optim/sched.py:13-16)
# ----------------------------------------------------------------------------
def no_decay(name):
    """optim/misc.py:11 substring rule."""
    return any(nd in name for nd in ("bias", "LayerNorm.bias", "LayerNorm.weight"))


def warmup_linear(step, warmup_step, tot_step):
    if step < warmup_step:
        return step / warmup_step
    return max(0, (tot_step - step) / (tot_step - warmup_step))


def clip_grad_norm(grads, max_norm):
    """torch.nn.utils.clip_grad_norm_ as called at pretrain.py:610."""
    total = torch.sqrt(sum((g.double() ** 2).sum() for g in grads)).float()
    coef = torch.clamp(max_norm / (total + 1e-6), max=1.0)
    for g in grads:
        g.mul_(coef)
    return total


def adamw_step(p, g, m, v, step, lr, beta1=0.9, beta2=0.98, eps=1e-6, weight_decay=0.0):
    """One parameter tensor, in place; ``step`` is the already-incremented count."""
    m.mul_(beta1).add_(g, alpha=1.0 - beta1)
    v.mul_(beta2).addcmul_(g, g, value=1.0 - beta2)
    denom = v.sqrt().add_(eps)
    step_size = lr * math.sqrt(1.0 - beta2 ** step) / (1.0 - beta1 ** step)
    p.addcdiv_(m, denom, value=-step_size)
    if weight_decay > 0.0:
        p.add_(p, alpha=-lr * weight_decay)
