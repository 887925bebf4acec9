"""Deterministic synthetic weights and batches for the UC2 cross-modal encoder path.

Everything here is pure numpy integer hashing (Irwin-Hall normals), so the GPU
box regenerates exactly the tensors the golden fixtures (tests/golden/) were
made from -- no dependence on torch's RNG stream or on CPU vector width.

Batch layouts restate what the reference collate functions produce
(/root/reference/data/itm.py:281-319 xlmr_itm_ot_collate,
 data/data.py:360-384 pad_tensors/get_gather_index,
 data/mrm.py:22-39, data/mlm.py:30-67); the index builders themselves live in
``uc2_b200.batch`` (product) and ``oracle/uc2_oracle.py`` (checker).
"""
import zlib

import numpy as np
import torch

IMG_DIM = 2048          # utils/const.py:2
IMG_LABEL_DIM = 1601    # utils/const.py:3

UC2_BASE = dict(        # config/uc2-base.json
    attention_probs_dropout_prob=0.1, hidden_act="gelu", hidden_dropout_prob=0.1,
    hidden_size=768, initializer_range=0.02, intermediate_size=3072,
    max_position_embeddings=514, num_attention_heads=12, num_hidden_layers=12,
    model_type="xlm-roberta", output_past=True, type_vocab_size=2,
    layer_norm_eps=1e-5, pad_token_id=1, vocab_size=250002)


# ----------------------------------------------------------------------------
# counter-based generator
# ----------------------------------------------------------------------------
def _mix64(x):
    """splitmix64 finaliser on a uint64 array (wraps mod 2^64)."""
    x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return x ^ (x >> np.uint64(31))


def det_u64(n, seed, stream=0):
    with np.errstate(over="ignore"):
        idx = np.arange(n, dtype=np.uint64)
        key = np.uint64((seed * 0x9E3779B97F4A7C15 + stream * 0xD1B54A32D192ED03) % (1 << 64))
        return _mix64(_mix64(idx + key) + np.uint64(0x9E3779B97F4A7C15))


def det_uniform(n, seed, stream=0):
    """float64 in (0,1), 53-bit."""
    return ((det_u64(n, seed, stream) >> np.uint64(11)).astype(np.float64) + 0.5) * (1.0 / (1 << 53))


def _normal_chunk(args):
    s, m, seed, std, mean = args
    with np.errstate(over="ignore"):
        idx = np.arange(s, s + m, dtype=np.uint64)
        k1 = np.uint64((seed * 0x9E3779B97F4A7C15 + 0xD1B54A32D192ED03) % (1 << 64))
        a = _mix64(_mix64(idx + k1) + np.uint64(0x9E3779B97F4A7C15))
    # Irwin-Hall(4) on the four 16-bit lanes of one hash: integer-exact, unit variance after
    # scaling by sqrt(3); close enough to N(0,1) for synthetic weights/features and, unlike
    # Box-Muller, free of libm calls (bit-identical on every host).
    m16 = np.uint64(0xFFFF)
    t = ((a & m16) + ((a >> np.uint64(16)) & m16) + ((a >> np.uint64(32)) & m16)
         + (a >> np.uint64(48))).astype(np.int64) - 131070
    return (np.float32(mean) + t.astype(np.float32) * np.float32(std * 1.7320508075688772 / 65536.0))


def det_normal(shape, seed, std=1.0, mean=0.0, chunk=1 << 22):
    n = int(np.prod(shape))
    jobs = [(s, min(chunk, n - s), seed, std, mean) for s in range(0, n, chunk)]
    if len(jobs) > 4:
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(8) as ex:
            parts = list(ex.map(_normal_chunk, jobs))
    else:
        parts = [_normal_chunk(j) for j in jobs]
    out = np.concatenate(parts) if len(parts) > 1 else parts[0]
    return out.reshape(shape)


def det_randint(n, lo, hi, seed, stream=0):
    """ints in [lo, hi)."""
    return (lo + (det_u64(n, seed, stream) % np.uint64(hi - lo)).astype(np.int64))


def _name_seed(name, seed):
    return (zlib.crc32(name.encode()) * 2654435761 + seed * 97) % (1 << 31)


# ----------------------------------------------------------------------------
# weights
# ----------------------------------------------------------------------------
def fill_state_dict(shapes, seed=42, perturb=True, std=0.02):
    """Deterministic values for a ``{name: shape}`` mapping.

    perturb=False restates the reference initialiser (model/model.py:159-172:
    Linear/Embedding weights N(0, 0.02), LayerNorm (1, 0), Linear bias 0).
    perturb=True (used by parity tests) also randomises biases and LayerNorm
    affine terms so that every term of every kernel is exercised.
    """
    out = {}
    for name, shape in shapes.items():
        s = _name_seed(name, seed)
        low = name.lower()
        is_ln = ("layernorm" in low or "layer_norm" in low or low.endswith("net.2.weight")
                 or low.endswith("net.2.bias"))
        if is_ln and name.endswith("weight"):
            v = det_normal(shape, s, 0.05, 1.0) if perturb else np.ones(shape, np.float32)
        elif name.endswith("bias") or (is_ln and name.endswith("bias")):
            v = det_normal(shape, s, std) if perturb else np.zeros(shape, np.float32)
        else:
            v = det_normal(shape, s, std)
        out[name] = torch.from_numpy(np.ascontiguousarray(v))
    return out


# ----------------------------------------------------------------------------
# batches
# ----------------------------------------------------------------------------
def _lens(n, lo, hi, seed, stream, fixed=None):
    if fixed is not None:
        return [int(fixed)] * n
    return [int(x) for x in det_randint(n, lo, hi + 1, seed, stream)]


def make_pairs(n, seed=42, txt_len=None, num_bb=None, txt_range=(8, 60), bb_range=(10, 100),
               vocab=250002, family="vlxlmr"):
    """Per-sample ragged inputs: list of dicts(input_ids[tl], img_feat[nbb,2048],
    img_pos_feat[nbb,7]). tl counts <s> and </s> (ids 0 and 2, data/data.py:216-220).
    Features are |N(0,1)| (ReLU-like Faster-RCNN feats); boxes are sorted
    U(0,1) corners extended to [x1,y1,x2,y2,w,h,w*h] (data/data.py:339)."""
    tls = _lens(n, txt_range[0], txt_range[1], seed, 11, txt_len)
    nbs = _lens(n, bb_range[0], bb_range[1], seed, 12, num_bb)
    cls_id, sep_id, lo_id = (0, 2, 5) if family == "vlxlmr" else (101, 102, 1000)
    items = []
    for i, (tl, nbb) in enumerate(zip(tls, nbs)):
        ids = det_randint(tl, lo_id, vocab, seed * 1000003 + i, 13)
        ids[0], ids[-1] = cls_id, sep_id
        feat = np.abs(det_normal((nbb, IMG_DIM), seed * 1000003 + i + (1 << 20)))
        c = det_uniform(nbb * 4, seed * 1000003 + i, 14).reshape(nbb, 4)
        x1, x2 = np.minimum(c[:, 0], c[:, 1]), np.maximum(c[:, 0], c[:, 1])
        y1, y2 = np.minimum(c[:, 2], c[:, 3]), np.maximum(c[:, 2], c[:, 3])
        w, h = x2 - x1, y2 - y1
        pos = np.stack([x1, y1, x2, y2, w, h, w * h], 1).astype(np.float32)
        items.append(dict(input_ids=torch.from_numpy(ids.astype(np.int64)),
                          img_feat=torch.from_numpy(feat), img_pos_feat=torch.from_numpy(pos)))
    return items


def make_mlm_labels(input_ids_list, seed, mask_id=250001, vocab=250002):
    """BERT-style 15% masking (data/mlm.py:30-67): returns masked ids + labels
    (-1 = not predicted); specials at both ends are never masked; >=1 masked."""
    outs = []
    for i, ids in enumerate(input_ids_list):
        ids = ids.clone()
        n = ids.numel() - 2
        u = det_uniform(n, seed * 7919 + i, 21)
        rnd = det_randint(n, 5, vocab, seed * 7919 + i, 22)
        lab = torch.full_like(ids, -1)
        any_masked = False
        for j in range(n):
            p = u[j]
            if p < 0.15:
                p /= 0.15
                lab[j + 1] = ids[j + 1]
                if p < 0.8:
                    ids[j + 1] = mask_id
                elif p < 0.9:
                    ids[j + 1] = int(rnd[j])
                any_masked = True
        if not any_masked:
            lab[1] = ids[1]
            ids[1] = mask_id
        outs.append((ids, lab))
    return outs


def make_img_masks(num_bbs, seed, prob=0.15):
    """data/mrm.py:13-19: Bernoulli(0.15) per region, at least one."""
    out = []
    for i, nbb in enumerate(num_bbs):
        u = det_uniform(nbb, seed * 104729 + i, 31)
        m = u < prob
        if not m.any():
            m[int(det_randint(1, 0, nbb, seed * 104729 + i, 32)[0])] = True
        out.append(torch.from_numpy(m))
    return out


def make_soft_labels(nbb, seed):
    """softmax(3*N(0,1)) over 1601 classes."""
    z = det_normal((nbb, IMG_LABEL_DIM), seed, 3.0).astype(np.float64)
    z = np.exp(z - z.max(1, keepdims=True))
    return torch.from_numpy((z / z.sum(1, keepdims=True)).astype(np.float32))
