"""Generate tests/golden/*.npz by running the UNMODIFIED reference modules from
/root/reference (with the import shims of ref_shims.py) on the deterministic cases of
tests/cases.py.  Runs only in the authoring container; the fixtures are committed and the
GPU box never needs /root/reference.

    python tests/golden/make_golden.py            # all cases
    python tests/golden/make_golden.py pretrain   # one case
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import ref_shims  # noqa: E402
import cases  # noqa: E402

R = ref_shims.load()
REF_CFG = "/root/reference/config/uc2-base.json"


def grad_digest(t):
    """[L2 norm, 48 strided samples] of a gradient tensor."""
    f = t.detach().reshape(-1).double()
    idx = torch.linspace(0, f.numel() - 1, 48).long()
    return np.concatenate([[f.norm().item()], f[idx].numpy()]).astype(np.float64)


def ref_config(cfg, family):
    C = R.model.VLXLMRConfig if family == "vlxlmr" else R.model.UniterConfig
    c = C.from_json_file(REF_CFG)
    c.__dict__.update(cfg.to_dict())
    return c


def build(kind, cfg, family, sd):
    rc = ref_config(cfg, family)
    if kind == "pretrain":
        M = R.model.VLXLMRForPretraining if family == "vlxlmr" else R.model.UniterForPretraining
        m = M(rc, 2048, 1601)
    else:
        M = R.itm.VLXLMRForImageTextRetrieval if family == "vlxlmr" else R.itm.UniterForImageTextRetrieval
        m = M(rc, 2048, margin=0.2)
    missing, unexpected = m.load_state_dict(cases.with_aliases(sd, kind, family), strict=False)
    assert not unexpected, unexpected
    assert all(k.startswith("vis_cls.") for k in missing), missing
    m.eval()   # dropout off (parity runs use p=0, utils/misc.py:54-60)
    return m


def named_unique(m):
    return [(n, p) for n, p in m.named_parameters() if not n.startswith("vis_cls.")]


def run_task(m, batch, task, out, tag, lam=0.1):
    m.zero_grad(set_to_none=True)
    res = m(batch, task=task, compute_loss=True)
    if task == "itm":
        itm, (pos, neg) = res
        out[f"{tag}|itm_loss"] = cases.to_np(itm)
        out[f"{tag}|ot_pos"] = cases.to_np(pos)
        out[f"{tag}|ot_neg"] = cases.to_np(neg)
        loss = itm.mean() + lam * (pos.sum() - neg.sum()) / (pos.size(0) + neg.size(0))   # pretrain.py:528-544
    else:
        out[f"{tag}|loss_vec"] = cases.to_np(res)
        loss = res.mean()
    out[f"{tag}|loss"] = np.array([loss.item()])
    loss.backward()
    for n, p in named_unique(m):
        if p.grad is not None:
            out[f"{tag}|grad|{n}"] = grad_digest(p.grad)
    with torch.no_grad():
        sc = m(batch, task=task, compute_loss=False)
        sc = sc[0] if task == "itm" else sc
        s = cases.to_np(sc)
        out[f"{tag}|scores"] = s if s.shape[-1] <= 2048 else s[:, ::97]


def hidden_dump(enc, fam, batch, out, tag, img_masks=None):
    with torch.no_grad():
        pos = batch["position_ids"] if fam == "uniter" else None
        emb = enc._compute_img_txt_embeddings(batch["input_ids"], pos, batch["img_feat"],
                                              batch["img_pos_feat"], batch["gather_index"], img_masks)
        layers = enc(batch["input_ids"], pos, batch["img_feat"], batch["img_pos_feat"],
                     batch["attn_masks"], batch["gather_index"], img_masks=img_masks,
                     output_all_encoded_layers=True)
    rows = cases.sample_rows(emb.size(1))
    out[f"{tag}|emb"] = cases.to_np(emb)            # full packed embedding (checks packing row by row)
    out[f"{tag}|hidden_rows"] = np.array(rows)
    out[f"{tag}|hidden"] = np.stack([cases.to_np(h[:, rows]) for h in layers])


def case_pretrain(family="vlxlmr"):
    cfg = cases.config(2, family=family)
    sd = cases.weights(cfg, "pretrain", family)
    m = build("pretrain", cfg, family, sd)
    enc = m.roberta if family == "vlxlmr" else m.bert
    out = {}
    b_itm = cases.batch_itm(family=family)
    hidden_dump(enc, family, b_itm, out, "itm")
    run_task(m, b_itm, "itm", out, "itm")
    run_task(m, cases.batch_mlm(family=family), "mlm", out, "mlm")
    b_mrfr = cases.batch_mrfr(family=family)
    hidden_dump(enc, family, b_mrfr, out, "mrfr", img_masks=b_mrfr["img_masks"])
    run_task(m, b_mrfr, "mrfr", out, "mrfr")
    b_mrc = cases.batch_mrc(family=family)
    run_task(m, b_mrc, "mrc-kl", out, "mrc-kl")
    run_task(m, b_mrc, "mrc", out, "mrc")
    if family == "vlxlmr":                       # MRTM tasks exist for the VLXLMR family only (model.py:522-543)
        run_task(m, cases.batch_mmxlm(), "mmxlm", out, "mmxlm")
        b_soft = dict(cases.batch_mmxlm_soft())
        b_soft.pop("valid_token_ids")           # the reference reads its module-level VALID_XLMR_TOKEN_IDS (shimmed)
        run_task(m, b_soft, "vmlm-soft", out, "vmlm-soft")
    return out


def case_rank(family="vlxlmr"):
    cfg = cases.config(2, family=family)
    sd = cases.weights(cfg, "retrieval", family)
    m = build("retrieval", cfg, family, sd)
    out = {}
    b = cases.batch_rank(family=family)
    m.zero_grad(set_to_none=True)
    loss = m(b, compute_loss=True)
    out["rank|loss_mat"] = cases.to_np(loss)
    loss.mean().backward()
    for n, p in named_unique(m):
        if p.grad is not None:
            out[f"rank|grad|{n}"] = grad_digest(p.grad)
    with torch.no_grad():
        out["rank|scores"] = cases.to_np(m(b, compute_loss=False))
    return out


def case_cfg1():
    """BASELINE.json configs[0]: uc2-base 12L, XLM-R vocab, B=8 x (40 tok + 36 regions), ITM forward."""
    cfg = cases.config(12, vocab=250002)
    sd = cases.weights(cfg, "retrieval")
    m = build("retrieval", cfg, "vlxlmr", sd)
    b = cases.batch_rank(n=8, sample_size=1, seed=42, vocab=250002, txt_len=40, num_bb=36)
    out = {}
    hidden_dump(m.roberta, "vlxlmr", b, out, "cfg1")
    out["cfg1|emb"] = out["cfg1|emb"][:, cases.sample_rows(76)]
    with torch.no_grad():
        out["cfg1|scores"] = cases.to_np(m(b, compute_loss=False))
    return out


def case_index():
    """Integer builders of the reference collates + position ids, on ragged lengths."""
    out = {}
    tls = [5, 17, 9, 30, 12, 8]
    nbs = [10, 36, 100, 11, 40, 23]
    T, S = max(tls), max(t + n for t, n in zip(tls, nbs))
    out["gather_index"] = R.data.get_gather_index(tls, nbs, len(tls), T, S).numpy()
    out["ot_scatter"] = R.data_itm._compute_ot_scatter(tls, T, S).numpy()
    out["txt_pad"] = R.data_itm._compute_pad(tls, T).numpy()
    out["img_pad"] = R.data_itm._compute_pad(nbs, max(nbs)).numpy()
    out["lens"] = np.array([tls, nbs])
    ids = cases.batch_mlm()["input_ids"]
    ids[2, 3] = 1      # a pad in the middle: positions skip it (model.py:288-290)
    out["pos_in"] = ids.numpy()
    out["pos_out"] = R.model.create_position_ids_from_input_ids(ids, 1).numpy()
    feats = [torch.arange(n * 3, dtype=torch.float32).view(n, 3) + i for i, n in enumerate(nbs)]
    out["pad_tensors"] = R.data.pad_tensors(feats, nbs).numpy()
    return out


def case_ot():
    out = {}
    b, m_, n_ = 4, 12, 20
    txt = torch.from_numpy(cases.synth.det_normal((b, m_, 768), 901))
    img = torch.from_numpy(cases.synth.det_normal((b, n_, 768), 902))
    tls, nbs = [12, 5, 9, 1], [20, 7, 13, 20]
    txt_pad = cases.B.compute_pad(tls, m_)
    img_pad = cases.B.compute_pad(nbs, n_)
    txt.requires_grad_(True)
    img.requires_grad_(True)
    d = R.ot.optimal_transport_dist(txt, img, txt_pad, img_pad)
    d.sum().backward()
    out["dist"] = cases.to_np(d)
    out["dtxt"] = cases.to_np(txt.grad)
    out["dimg"] = cases.to_np(img.grad)
    out["lens"] = np.array([tls, nbs])
    return out


def case_adamw():
    """optim/adamw.py + optim/misc.py grouping + clip_grad_norm_ (pretrain.py:610), 4 steps."""
    out = {}
    names = ["enc.dense.weight", "enc.dense.bias", "enc.LayerNorm.weight", "img_layer_norm.weight"]
    shapes = [(33, 17), (17,), (17,), (17,)]

    class Mod(torch.nn.Module):
        pass
    mod = Mod()
    for i, (n, s) in enumerate(zip(names, shapes)):
        mod.register_parameter(n.replace(".", "_DOT_"), torch.nn.Parameter(
            torch.from_numpy(cases.synth.det_normal(s, 700 + i, 0.5))))
    plist = [(n.replace("_DOT_", "."), p) for n, p in mod.named_parameters()]
    mod.named_parameters = lambda: plist

    class Opts:
        weight_decay, optim, learning_rate, betas = 0.01, "adamw", 3e-3, (0.9, 0.98)
    opt = R.misc.build_optimizer(mod, Opts)
    out["decay_flags"] = np.array([len(g["params"]) for g in opt.param_groups])
    import warnings
    warnings.simplefilter("ignore")
    for step in range(1, 5):
        lr = Opts.learning_rate * R.sched.warmup_linear(step, 2, 10)
        for g in opt.param_groups:
            g["lr"] = lr
        for i, (n, p) in enumerate(mod.named_parameters()):
            p.grad = torch.from_numpy(cases.synth.det_normal(p.shape, 800 + 10 * step + i, 2.0))
        gn = torch.nn.utils.clip_grad_norm_([p for _, p in plist], 5.0)
        out[f"gnorm{step}"] = np.array([float(gn)])
        opt.step()
        for n, p in mod.named_parameters():
            out[f"p{step}|{n}"] = cases.to_np(p)
    return out


def case_retrieval():
    """Retrieval scoring + recall: the reference model scores every (caption, image) pair through batches built
    the way ItmValDataset.get_batch does (data/itm.py:456-485: one caption x a chunk of images sorted by box
    count, reference pad_tensors / get_gather_index), stored as the fp16 score matrix of itm.py:515-538, and the
    reference's own itm_eval (eval/itm.py) turns it -- and a larger random matrix -- into recall numbers."""
    out = {}
    cfg = cases.config(2)
    sd = cases.weights(cfg, "retrieval")
    m = build("retrieval", cfg, "vlxlmr", sd)
    images, captions, txt_ids, txt2img, img2txts = cases.retrieval_case()
    order = sorted(range(len(images)), key=lambda i: images[i]["img_feat"].size(0))
    img_ids = [images[i]["id"] for i in order]
    bs = 4
    score = torch.zeros(len(captions), len(images), dtype=torch.float16)
    with torch.no_grad():
        for ci, ids in enumerate(captions):
            j = 0
            for st in range(0, len(order), bs):
                sel = order[st:st + bs]
                nbs = [images[i]["img_feat"].size(0) for i in sel]
                tl = ids.numel()
                input_ids = ids.unsqueeze(0).expand(len(sel), -1).clone()
                img_feat = R.data.pad_tensors([images[i]["img_feat"] for i in sel], nbs)
                img_pos = R.data.pad_tensors([images[i]["img_pos_feat"] for i in sel], nbs)
                attn = torch.zeros(len(sel), max(nbs) + tl).long()
                for k, nb in enumerate(nbs):
                    attn.data[k, :tl + nb].fill_(1)
                gi = R.data.get_gather_index([tl] * len(sel), nbs, len(sel), tl, attn.size(1))
                batch = dict(input_ids=input_ids, position_ids=None, img_feat=img_feat, img_pos_feat=img_pos,
                             attn_masks=attn, gather_index=gi)
                sc = m(batch, compute_loss=False)
                score[ci, j:j + len(sel)] = sc.squeeze(1).half()
                j += len(sel)
    out["retrieval|scores"] = score.float().numpy()
    out["retrieval|img_order"] = np.array(order)
    log = R.eval_itm.itm_eval(score.float(), txt_ids, img_ids, txt2img, img2txts)
    out["retrieval|recall"] = np.array([log[k] for k in sorted(log)])
    # recall on a larger random matrix (60 captions x 15 images, 4 captions per image)
    n_img, cpi = 15, 4
    big = torch.from_numpy(cases.synth.det_normal((n_img * cpi, n_img), 4242)).float()
    tids = [f"t{i}" for i in range(n_img * cpi)]
    iids = [f"i{j}" for j in range(n_img)]
    t2i = {t: f"i{i // cpi}" for i, t in enumerate(tids)}
    i2t = {f"i{j}": [tids[j * cpi + k] for k in range(cpi)] for j in range(n_img)}
    for j in range(n_img):                      # make the ground truth partly retrievable
        for k in range(cpi):
            big[j * cpi + k, j] += 1.5 * ((j + k) % 3)
    log2 = R.eval_itm.itm_eval(big, tids, iids, t2i, i2t)
    out["recall|matrix"] = big.numpy()
    out["recall|values"] = np.array([log2[k] for k in sorted(log2)])
    out["recall|keys"] = np.array(sorted(log2))
    return out


def case_loader():
    """data/sampler.py TokenBucketSampler and data/loader.py MetaLoader driven by the seeded `random` module."""
    import random
    out = {}
    lens = [int(x) for x in cases.synth.det_randint(700, 18, 161, 5, 1)]
    out["lens"] = np.array(lens)
    for k, (bucket, budget, drop) in enumerate([(256, 2560, False), (128, 10240, True), (512, 1920, False)]):
        random.seed(100 + k)
        batches = [b for b in iter(R.sampler.TokenBucketSampler(lens, bucket, budget, droplast=drop))]
        out[f"sampler{k}|cfg"] = np.array([bucket, budget, int(drop)])
        out[f"sampler{k}|sizes"] = np.array([len(b) for b in batches])
        out[f"sampler{k}|flat"] = np.array([i for b in batches for i in b])

    class Loader(torch.utils.data.DataLoader):      # MetaLoader insists on DataLoader instances
        def __init__(self, items):
            self.items = items

        def __iter__(self):
            return iter(self.items)
    random.seed(7)
    ml = R.loader.MetaLoader({"mlm": (Loader([1, 2, 3]), 2), "itm": Loader([10, 20]), "mrfr": (Loader([5]), 1)},
                             accum_steps=3, distributed=False)
    seq = []
    for i, (task, batch) in enumerate(ml):
        seq.append((task, batch))
        if i == 59:
            break
    out["meta|tasks"] = np.array([t for t, _ in seq])
    out["meta|batches"] = np.array([b for _, b in seq])
    return out


SCHED_OPTS = [dict(decay=d, learning_rate=3e-4, xlmr_lr=1e-5, warmup_steps=10, num_train_steps=50, warm_int=4,
                   decay_int=7, decay_st=20, decay_rate=0.5) for d in ("linear", "invsqrt", "constant", "vqa")]


def case_sched():
    """optim/sched.py get_lr_sched / get_xlmr_lr_sched for steps 0..59 under every decay mode (incl. the <= 0 guard)."""
    import types
    out = {}
    for o in SCHED_OPTS:
        ns = types.SimpleNamespace(**o)
        out[f"{o['decay']}|lr"] = np.array([R.sched.get_lr_sched(s, ns) for s in range(60)], dtype=np.float64)
        out[f"{o['decay']}|xlmr_lr"] = np.array([R.sched.get_xlmr_lr_sched(s, ns) for s in range(60)], dtype=np.float64)
    return out


def _digest(t):
    """Small fingerprint of a float tensor: shape, per-sample sums (fp64) and 64 strided samples."""
    f = t.detach().double()
    flat = f.reshape(-1)
    idx = torch.linspace(0, flat.numel() - 1, 64).long()
    return np.concatenate([np.array(f.shape, dtype=np.float64), f.reshape(f.size(0), -1).sum(1).numpy(),
                           flat[idx].numpy()])


def _dump_batch(out, tag, batch):
    for k, v in batch.items():
        if isinstance(v, dict):
            _dump_batch(out, f"{tag}|{k}", v)
        elif torch.is_tensor(v):
            out[f"{tag}|{k}"] = _digest(v) if v.is_floating_point() else v.numpy().astype(np.int64)
        else:
            out[f"{tag}|{k}"] = np.array([v])


def case_collate():
    """The reference's WHOLE collate functions (data/itm.py, data/mrm.py, data/mlm.py) on the deterministic items of
    tests/cases.py: integer / mask outputs stored in full, float tensors as digests."""
    import importlib
    cwd = os.getcwd()
    os.chdir("/root/reference")
    try:
        mlm = importlib.import_module("data.mlm")
        mrm = importlib.import_module("data.mrm")
        ditm = importlib.import_module("data.itm")
    finally:
        os.chdir(cwd)
    out = {}
    items = cases._items(6, 71, cases.SMALL_VOCAB, "vlxlmr")
    ones = lambda it: torch.ones(it["input_ids"].numel() + it["img_feat"].size(0), dtype=torch.long)
    targets = [1, 0, 0, 1, 1, 0]
    _dump_batch(out, "itm_ot", ditm.xlmr_itm_ot_collate(
        [(it["input_ids"], it["img_feat"], it["img_pos_feat"], ones(it), torch.tensor([t]))
         for it, t in zip(items, targets)]))
    _dump_batch(out, "itm", ditm.xlmr_itm_collate(
        [(it["input_ids"], it["img_feat"], it["img_pos_feat"], ones(it), torch.tensor([t]))
         for it, t in zip(items, targets)]))
    _dump_batch(out, "rank", ditm.xlmr_itm_rank_collate(
        [[(it["input_ids"], it["img_feat"], it["img_pos_feat"], ones(it)) for it in items]]))
    lab = cases.synth.make_mlm_labels([it["input_ids"] for it in items], 71, mask_id=cases.SMALL_VOCAB - 1,
                                      vocab=cases.SMALL_VOCAB)
    _dump_batch(out, "mlm", mlm.xlmr_mlm_collate(
        [(m, it["img_feat"], it["img_pos_feat"], ones(it), l) for it, (m, l) in zip(items, lab)]))
    nbbs = [it["img_feat"].size(0) for it in items]
    masks = cases.synth.make_img_masks(nbbs, 71)
    tgt = [mrm._get_img_tgt_mask(mk, it["input_ids"].numel()) for mk, it in zip(masks, items)]
    _dump_batch(out, "mrfr", mrm.xlmr_mrfr_collate(
        [(it["input_ids"], it["img_feat"], it["img_pos_feat"], ones(it), mk, tg)
         for it, mk, tg in zip(items, masks, tgt)]))
    # VTLM with images (co-masking datasets): per-sample position ids travel through the collate
    from uc2_b200.batch import tlm_position_ids
    _dump_batch(out, "tlm", mlm.xlmr_mlm_dmasking_collate(
        [(m, it["img_feat"], it["img_pos_feat"], ones(it), l, tlm_position_ids(m)) for it, (m, l) in zip(items, lab)]))
    # MRTM: hard token labels over the packed sequence / soft token distributions of the masked regions
    img_lab = []
    for i, (nb, mk) in enumerate(zip(nbbs, masks)):
        tok = torch.from_numpy(cases.synth.det_randint(nb, 5, cases.SMALL_VOCAB, 71 * 77 + i, 5).astype(np.int64))
        img_lab.append(torch.where(mk, tok, torch.full_like(tok, -1)))
    _dump_batch(out, "mmxlm", mlm.xlmr_mmxlm_collate(
        [(m, it["img_feat"], it["img_pos_feat"], ones(it), mk, torch.cat([l, il]))
         for it, (m, l), mk, il in zip(items, lab, masks, img_lab)]))
    tok_soft = [torch.softmax(torch.from_numpy(cases.synth.det_normal((nb, 16), 71 * 5 + i, 2.0)).float(), -1)
                for i, nb in enumerate(nbbs)]
    _dump_batch(out, "mmxlm_soft", mlm.xlmr_mmxlm_softlabel_collate(
        [(it["input_ids"], it["img_feat"], it["img_pos_feat"], ones(it), mk, tg, ts)
         for it, mk, tg, ts in zip(items, masks, tgt, tok_soft)]))
    soft = [cases.synth.make_soft_labels(nb, 71 * 31 + i) for i, nb in enumerate(nbbs)]
    _dump_batch(out, "mrc", mrm.xlmr_mrc_collate(
        [(it["input_ids"], it["img_feat"], it["img_pos_feat"], sl, ones(it), mk, tg)
         for it, sl, mk, tg in zip(items, soft, masks, tgt)]))
    return out


def case_sampling():
    """data/mlm.py random_word, data/mrm.py _get_img_mask, data/itm.py sample_negative + the id pairs of
    ItmRankDataset.__getitem__, all under fixed `random` seeds."""
    import random
    import importlib
    cwd = os.getcwd()
    os.chdir("/root/reference")                 # data/mlm.py reads object_labels/*.txt relative to the cwd at import
    try:
        mlm = importlib.import_module("data.mlm")
        mrm = importlib.import_module("data.mrm")
        ditm = importlib.import_module("data.itm")
    finally:
        os.chdir(cwd)
    out = {}
    lens = [1, 3, 9, 17, 40, 58]
    random.seed(11)
    toks, labs = [], []
    for k, n in enumerate(lens * 3):
        ids = [int(x) for x in cases.synth.det_randint(n, 5, 250001, 300 + k, 2)]
        t, l = mlm.random_word(ids, (5, 250001), 250001)
        toks += t
        labs += l
    out["random_word|lens"] = np.array(lens * 3)
    out["random_word|tokens"] = np.array(toks)
    out["random_word|labels"] = np.array(labs)
    random.seed(12)
    nbbs = [1, 2, 10, 36, 100, 5, 3, 64]
    out["img_mask|nbbs"] = np.array(nbbs)
    out["img_mask|flat"] = np.concatenate([mrm._get_img_mask(0.15, n).numpy().astype(np.uint8) for n in nbbs])
    random.seed(13)
    imgs = [f"img{i}" for i in range(12)]
    txts = [f"txt{i}" for i in range(36)]
    img2txts = {f"img{i}": [f"txt{3 * i + k}" for k in range(3)] for i in range(12)}
    pairs = []
    for t in (0, 7, 20, 35, 14):
        gi = f"img{t // 3}"
        for ns in (1, 2):
            neg_i = ditm.sample_negative(imgs, [gi], ns)
            neg_t = ditm.sample_negative(txts, img2txts[gi], ns)
            pairs += [f"txt{t}|{gi}"] + [f"txt{t}|{i}" for i in neg_i] + [f"{x}|{gi}" for x in neg_t]
    out["rank|pairs"] = np.array(pairs)
    return out


CASES = {
    "collate": case_collate,
    "sampling": case_sampling,
    "sched": case_sched,
    "loader": case_loader,
    "retrieval": case_retrieval,
    "pretrain": lambda: case_pretrain("vlxlmr"),
    "pretrain_uniter": lambda: case_pretrain("uniter"),
    "rank": lambda: case_rank("vlxlmr"),
    "cfg1": case_cfg1,
    "index": case_index,
    "ot": case_ot,
    "adamw": case_adamw,
}

if __name__ == "__main__":
    torch.manual_seed(0)
    torch.set_num_threads(8)
    which = sys.argv[1:] or list(CASES)
    for name in which:
        res = CASES[name]()
        path = os.path.join(HERE, f"{name}.npz")
        np.savez_compressed(path, **res)
        print(f"{name}: {len(res)} arrays -> {path} ({os.path.getsize(path) / 1e3:.0f} kB)")
