"""GPU: device-side batch assembly (uc2_b200/device_batch.py over uc2_pad_rows / uc2_batch_index) against the
host collates of uc2_b200/batch.py, which tests/test_oracle_golden.py pins to the reference's own collate
functions (data/data.py, data/itm.py, data/mrm.py, data/mlm.py).  Copies and integer arithmetic: bit-exact."""
import numpy as np
import pytest
import torch

import cases
from uc2_b200 import synth

pytestmark = pytest.mark.gpu


def _same(got, want, path=""):
    assert set(got) == set(want), (path, sorted(got), sorted(want))
    for k, w in want.items():
        g = got[k]
        if isinstance(w, dict):
            _same(g, w, path + k + ".")
        elif torch.is_tensor(w):
            assert torch.is_tensor(g) and g.is_cuda, path + k
            assert g.dtype == w.dtype and tuple(g.shape) == tuple(w.shape), (path + k, g.dtype, w.dtype, g.shape, w.shape)
            assert torch.equal(g.cpu(), w), path + k
        else:
            assert g == w, (path + k, g, w)


def _setup(n_img=9, n=7, seed=3, **kw):
    from uc2_b200.device_batch import DeviceCollator, FeatureArena
    imgs = cases._items(n_img, seed, cases.SMALL_VOCAB, "vlxlmr", **kw)
    txts = cases._items(n, seed + 1, cases.SMALL_VOCAB, "vlxlmr", **kw)
    soft = [synth.make_soft_labels(it["img_feat"].size(0), seed * 31 + i) for i, it in enumerate(imgs)]
    arena = FeatureArena([it["img_feat"] for it in imgs], [it["img_pos_feat"] for it in imgs], soft)
    img_idx = [int(x) for x in synth.det_randint(n, 0, n_img, seed, 17)]          # repeats allowed
    items = [dict(input_ids=t["input_ids"], img_feat=imgs[i]["img_feat"], img_pos_feat=imgs[i]["img_pos_feat"])
             for t, i in zip(txts, img_idx)]
    return DeviceCollator(arena), items, img_idx, [soft[i] for i in img_idx]


@pytest.mark.parametrize("kw", [dict(), dict(txt_len=60, num_bb=100), dict(bb_range=(1, 3))])
def test_itm_and_rank(kw):
    from uc2_b200 import batch as B
    dc, items, img_idx, _ = _setup(**kw)
    ids = [it["input_ids"] for it in items]
    targets = [i % 2 for i in range(len(items))]
    for with_ot in (True, False):
        _same(dc.itm(ids, img_idx, targets, with_ot=with_ot), B.collate_itm(items, targets, with_ot=with_ot))
    dc6, items6, idx6, _ = _setup(n=6, **kw)
    _same(dc6.itm_rank([it["input_ids"] for it in items6], idx6, 3), B.collate_itm_rank(items6, 3))


def test_mlm_and_tlm():
    from uc2_b200 import batch as B
    dc, items, img_idx, _ = _setup()
    lab = synth.make_mlm_labels([it["input_ids"] for it in items], 5, mask_id=cases.SMALL_VOCAB - 1,
                                vocab=cases.SMALL_VOCAB)
    _same(dc.mlm(lab, img_idx), B.collate_mlm(items, lab))
    for it in items:                       # a second <s> so the TLM positions restart
        it["input_ids"][it["input_ids"].numel() // 2] = 0
    lab = synth.make_mlm_labels([it["input_ids"] for it in items], 6, mask_id=cases.SMALL_VOCAB - 1,
                                vocab=cases.SMALL_VOCAB)
    _same(dc.tlm(lab, img_idx), B.collate_tlm(items, lab))


@pytest.mark.parametrize("seed", [3, 11])
def test_mrfr_mrc_mmxlm(seed):
    from uc2_b200 import batch as B
    dc, items, img_idx, soft = _setup(seed=seed)
    nbbs = [it["img_feat"].size(0) for it in items]
    masks = synth.make_img_masks(nbbs, seed)
    ids = [it["input_ids"] for it in items]
    _same(dc.mrfr(ids, img_idx, masks), B.collate_mrfr(items, masks))
    _same(dc.mrc(ids, img_idx, masks), B.collate_mrc(items, masks, soft))
    lab = synth.make_mlm_labels(ids, seed, mask_id=cases.SMALL_VOCAB - 1, vocab=cases.SMALL_VOCAB)
    img_lab = []
    for i, (nb, mk) in enumerate(zip(nbbs, masks)):
        tok = torch.from_numpy(synth.det_randint(nb, 5, cases.SMALL_VOCAB, seed * 77 + i, 5).astype(np.int64))
        img_lab.append(torch.where(mk, tok, torch.full_like(tok, -1)))
    _same(dc.mmxlm(lab, img_idx, masks, img_lab), B.collate_mmxlm(items, lab, masks, img_lab))


def test_model_consumes_device_batch():
    """The assembled batch drives the model to the same loss as the host-collated one."""
    from uc2_b200 import batch as B, model
    from uc2_b200.utils import set_dropout
    dc, items, img_idx, _ = _setup(seed=5)
    cfg = cases.config(2)
    m = model.VLXLMRForPretraining(cfg, 2048, 1601)
    m.load_state_dict(cases.with_aliases(cases.weights(cfg, "pretrain"), "pretrain"), strict=False)
    m.cuda().eval()
    set_dropout(m, 0)
    masks = synth.make_img_masks([it["img_feat"].size(0) for it in items], 5)
    ids = [it["input_ids"] for it in items]
    with torch.no_grad():
        a = m(dc.mrfr(ids, img_idx, masks), task="mrfr", compute_loss=True)
        b = m(B.to_device(B.collate_mrfr(items, masks), "cuda"), task="mrfr", compute_loss=True)
    torch.testing.assert_close(a, b, rtol=1e-6, atol=0)


def test_pipeline_from_stores_to_training_loop(tmp_path):
    """TextDB + FeatureArena -> datasets -> TokenBucketSampler -> DeviceCollator -> MetaLoader -> PretrainLoop:
    the whole caller side of the path, three tasks, two optimizer steps on a 2-layer model."""
    import random
    from types import SimpleNamespace
    from uc2_b200 import datasets as DS, model
    from uc2_b200.device_batch import DeviceCollator, FeatureArena
    from uc2_b200.loader import MetaLoader, TokenBucketSampler
    from uc2_b200.optim import AdamW
    from uc2_b200.pretrain_loop import PretrainLoop
    from uc2_b200.utils import set_dropout
    V = cases.SMALL_VOCAB
    imgs = cases._items(12, 41, V, "vlxlmr", bb_range=(10, 30))
    soft = [synth.make_soft_labels(it["img_feat"].size(0), 500 + i) for i, it in enumerate(imgs)]
    arena = FeatureArena([it["img_feat"] for it in imgs], [it["img_pos_feat"] for it in imgs], soft)
    names = [f"img{i}" for i in range(12)]
    caps = cases._items(36, 42, V, "vlxlmr", txt_range=(6, 20))
    ex = {f"t{k}": {"input_ids": c["input_ids"][1:-1].tolist(), "img_fname": f"img{k // 3}"} for k, c in enumerate(caps)}
    db = DS.TextDB(ex, mask=V - 1, v_range=(5, V - 1))
    idx = DS.ImageIndex(arena, names)
    dc = DeviceCollator(arena)
    random.seed(1); np.random.seed(1)
    dsets = {"mlm_coco": DS.MlmDataset(db, idx), "mrfr_coco": DS.MrfrDataset(0.15, db, idx),
             "mrc-kl_coco": DS.MrcDataset(0.15, db, idx), "itm_coco": DS.ItmDataset(db, idx)}
    loaders = {}
    for name, d in dsets.items():
        sampler = TokenBucketSampler(d.lens, 16, 8 * 60, droplast=True)
        loaders[name] = DS.BatchLoader(d, sampler, lambda items, d=d: type(d).collate(dc, items))
    b = next(iter(loaders["mrc-kl_coco"]))
    assert b["img_feat"].is_cuda and b["label_targets"].shape[1] == 1601 and b["input_ids"].shape[0] % 8 == 0
    cfg = cases.config(2)
    m = model.VLXLMRForPretraining(cfg, 2048, 1601)
    m.load_state_dict(cases.with_aliases(cases.weights(cfg, "pretrain"), "pretrain"), strict=False)
    m.cuda().train()
    set_dropout(m, 0.1)
    opt = AdamW([{"params": list(m.parameters()), "weight_decay": 0.01}], lr=1e-4, betas=(0.9, 0.98))
    opts = SimpleNamespace(gradient_accumulation_steps=1, num_train_steps=4, valid_steps=10, grad_norm=5.0,
                           itm_ot_lambda=0.1, ot_pos_only=False, learning_rate=1e-4, decay="linear", warmup_steps=1)
    loop = PretrainLoop(m, opt, opts, log_every=2)
    assert loop.run(MetaLoader(loaders)) == 4
    seen = [k for k, mt in loop.task2loss.items() if mt.val is not None]
    assert seen and all(np.isfinite(loop.task2loss[k].val) for k in seen)
    assert sum(loop.n_examples.values()) >= 4 * 8
