"""Deterministic parity cases shared by tests/golden/make_golden.py (which runs the
reference on them), the oracle tests (CPU) and the CUDA parity tests (GPU)."""
import numpy as np
import torch

from uc2_b200 import batch as B
from uc2_b200 import synth
from uc2_b200.config import UC2Config, pretraining_shapes, retrieval_shapes

SMALL_VOCAB = 8192


def config(layers=2, vocab=SMALL_VOCAB, family="vlxlmr"):
    cfg = UC2Config(num_hidden_layers=layers, vocab_size=vocab)
    if family == "uniter":
        cfg.max_position_embeddings = 512
        cfg.pad_token_id = 0
    return cfg


def weights(cfg, kind, family="vlxlmr", seed=42):
    shapes = pretraining_shapes(cfg, family) if kind == "pretrain" else retrieval_shapes(cfg, family)
    sd = synth.fill_state_dict(shapes, seed=seed, perturb=True)
    return sd


def with_aliases(sd, kind, family="vlxlmr"):
    """Add the tied aliases a reference state_dict carries."""
    sd = dict(sd)
    enc = "roberta." if family == "vlxlmr" else "bert."
    if kind == "pretrain":
        W = sd[enc + "embeddings.word_embeddings.weight"]
        if family == "vlxlmr":
            sd["cls.decoder.weight"] = W
            sd["cls.decoder.bias"] = sd["cls.bias"]
        else:
            sd["cls.predictions.decoder.weight"] = W
        sd["feat_regress.weight"] = sd[enc + "img_embeddings.img_linear.weight"]
    return sd


def _items(n, seed, vocab, family, txt_range=(8, 30), bb_range=(10, 40), txt_len=None, num_bb=None):
    return synth.make_pairs(n, seed=seed, txt_len=txt_len, num_bb=num_bb, txt_range=txt_range,
                            bb_range=bb_range, vocab=vocab, family=family)


def pad_id(family):
    return 1 if family == "vlxlmr" else 0


def batch_itm(n=6, seed=7, vocab=SMALL_VOCAB, family="vlxlmr", with_ot=True, **kw):
    items = _items(n, seed, vocab, family, **kw)
    targets = [int(x) for x in synth.det_randint(n, 0, 2, seed, 41)]
    if with_ot and len(set(targets)) == 1:
        targets[0] = 1 - targets[0]
    return B.collate_itm(items, targets, with_ot=with_ot, pad_id=pad_id(family))


def batch_mlm(n=6, seed=8, vocab=SMALL_VOCAB, family="vlxlmr", **kw):
    items = _items(n, seed, vocab, family, **kw)
    lab = synth.make_mlm_labels([it["input_ids"] for it in items], seed, mask_id=vocab - 1, vocab=vocab)
    return B.collate_mlm(items, lab, pad_id=pad_id(family))


def batch_mrfr(n=6, seed=9, vocab=SMALL_VOCAB, family="vlxlmr", **kw):
    items = _items(n, seed, vocab, family, **kw)
    masks = synth.make_img_masks([it["img_feat"].size(0) for it in items], seed)
    return B.collate_mrfr(items, masks, pad_id=pad_id(family))


def batch_mrc(n=6, seed=10, vocab=SMALL_VOCAB, family="vlxlmr", **kw):
    items = _items(n, seed, vocab, family, **kw)
    nbbs = [it["img_feat"].size(0) for it in items]
    masks = synth.make_img_masks(nbbs, seed)
    soft = [synth.make_soft_labels(nb, seed * 31 + i) for i, nb in enumerate(nbbs)]
    return B.collate_mrc(items, masks, soft, pad_id=pad_id(family))


def batch_rank(n=6, sample_size=3, seed=11, vocab=SMALL_VOCAB, family="vlxlmr", **kw):
    items = _items(n, seed, vocab, family, **kw)
    return B.collate_itm_rank(items, sample_size, pad_id=pad_id(family))


def sample_rows(S):
    """Row subset of a [B,S,768] activation stored in the fixtures."""
    return sorted(set([0, 1, S // 3, S // 2, S - 2, S - 1]))


def to_np(x):
    return x.detach().cpu().float().numpy().copy() if torch.is_tensor(x) else np.asarray(x)


def retrieval_case(n_img=10, caps_per_img=2, seed=21, vocab=SMALL_VOCAB, family="vlxlmr"):
    """A small retrieval evaluation set: images with ragged box counts, `caps_per_img` captions each
    (caption c belongs to image c // caps_per_img), as ItmEvalDataset presents them (data/itm.py:891-902)."""
    imgs = _items(n_img, seed, vocab, family, txt_range=(4, 4), bb_range=(10, 36))
    caps = _items(n_img * caps_per_img, seed + 1, vocab, family, txt_range=(5, 16), bb_range=(10, 10))
    images = [dict(img_feat=it["img_feat"], img_pos_feat=it["img_pos_feat"], id=f"img{i}") for i, it in enumerate(imgs)]
    captions = [it["input_ids"] for it in caps]
    txt_ids = [f"txt{i}" for i in range(len(captions))]
    txt2img = {t: f"img{i // caps_per_img}" for i, t in enumerate(txt_ids)}
    img2txts = {f"img{j}": [txt_ids[j * caps_per_img + k] for k in range(caps_per_img)] for j in range(n_img)}
    return images, captions, txt_ids, txt2img, img2txts


def batch_tlm(n=4, seed=12, vocab=SMALL_VOCAB, half_len=60, num_bb=100):
    """BASELINE.json configs[4] (VTLM): <s> src </s> <s> tgt </s> with positions restarting at the second <s>
    (data/mlm.py:420-428), both halves masked, + 100 regions: S = 2 * half_len + 2 + num_bb."""
    items = _items(n, seed, vocab, "vlxlmr", txt_len=2 * half_len + 2, num_bb=num_bb)
    for it in items:
        ids = it["input_ids"]
        ids[half_len] = 2          # </s> closing the source half
        ids[half_len + 1] = 0      # <s> opening the target half
    lab = synth.make_mlm_labels([it["input_ids"] for it in items], seed, mask_id=vocab - 1, vocab=vocab)
    lab = [(m, l) for m, l in lab]
    for (m, l), it in zip(lab, items):                      # never mask the inner specials
        for k in (half_len, half_len + 1):
            m[k] = it["input_ids"][k]
            l[k] = -1
    return B.collate_tlm(items, lab)


VALID_TOKEN_IDS = list(range(16))       # what tests/golden/ref_shims.py installs as VALID_XLMR_TOKEN_IDS


def batch_mmxlm(n=6, seed=13, vocab=SMALL_VOCAB, family="vlxlmr", **kw):
    """MRTM hard labels: masked words AND masked regions predict token ids (data/mlm.py:440-468)."""
    items = _items(n, seed, vocab, family, **kw)
    lab = synth.make_mlm_labels([it["input_ids"] for it in items], seed, mask_id=vocab - 1, vocab=vocab)
    nbbs = [it["img_feat"].size(0) for it in items]
    masks = synth.make_img_masks(nbbs, seed)
    img_lab = []
    for i, (nb, mk) in enumerate(zip(nbbs, masks)):
        tok = torch.from_numpy(synth.det_randint(nb, 5, vocab, seed * 77 + i, 5).astype(np.int64))
        img_lab.append(torch.where(mk, tok, torch.full_like(tok, -1)))
    return B.collate_mmxlm(items, lab, masks, img_lab, pad_id=pad_id(family))


def batch_mmxlm_soft(n=6, seed=14, vocab=SMALL_VOCAB, family="vlxlmr", **kw):
    """MRTM soft labels: masked regions predict a distribution over the valid token ids (data/mlm.py:319-345)."""
    items = _items(n, seed, vocab, family, **kw)
    nbbs = [it["img_feat"].size(0) for it in items]
    masks = synth.make_img_masks(nbbs, seed)
    soft = []
    for i, nb in enumerate(nbbs):
        z = torch.from_numpy(synth.det_normal((nb, len(VALID_TOKEN_IDS)), seed * 91 + i, 2.0)).float()
        soft.append(torch.softmax(z, -1))
    b = B.collate_mmxlm_soft(items, masks, soft, pad_id=pad_id(family))
    b["valid_token_ids"] = torch.tensor(VALID_TOKEN_IDS)
    return b
