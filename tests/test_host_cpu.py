"""CPU-only checks: the C-ABI library loads and exports every symbol include/uc2_b200.h declares, the
product path refuses to run without a GPU, host-side module/state_dict contracts, and the world_size-2
data-parallel logic on gloo."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

import cases

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _lib_path():
    from uc2_b200 import build
    return build.build()


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(_lib_path())
    header = open(os.path.join(ROOT, "include", "uc2_b200.h")).read()
    declared = set(re.findall(r"UC2_API\s+[\w\s\*]+?\b(uc2_\w+)\s*\(", header))
    assert len(declared) >= 30, declared
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/uc2_b200.h but not exported"
    from uc2_b200 import _lib
    assert set(_lib.EXPORTS) <= declared, set(_lib.EXPORTS) - declared
    lib.uc2_version.restype = ctypes.c_int
    assert lib.uc2_version() >= 100


def test_runtime_switches_are_host_state():
    """uc2_gemm_sched_dynamic / uc2_reserve_sms only set host-side state (no device needed) and hand back the previous
    value; test-support entry points reject bad arguments before touching a device."""
    lib = ctypes.CDLL(_lib_path())
    for f in (lib.uc2_gemm_sched_dynamic, lib.uc2_reserve_sms):
        f.restype, f.argtypes = ctypes.c_int, [ctypes.c_int]
    prev = lib.uc2_gemm_sched_dynamic(1)
    assert prev in (0, 1)
    assert lib.uc2_gemm_sched_dynamic(0) == 1
    assert lib.uc2_gemm_sched_dynamic(prev) == 0
    r0 = lib.uc2_reserve_sms(8)
    assert lib.uc2_reserve_sms(7) == 8          # odd requests are rounded down to CTA pairs
    assert lib.uc2_reserve_sms(r0) == 6
    lib.uc2_debug_occupy_sms.restype = ctypes.c_int
    lib.uc2_debug_occupy_sms.argtypes = [ctypes.c_int, ctypes.c_longlong, ctypes.c_void_p]
    assert lib.uc2_debug_occupy_sms(0, 10, None) < 0
    assert lib.uc2_debug_occupy_sms(4, -1, None) < 0


def test_no_gpu_means_loud_failure():
    """No CPU fallback: a compute entry point on a box without a usable sm_100 device returns an error code,
    and the Python modules raise."""
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    lib = ctypes.CDLL(_lib_path())
    lib.uc2_last_error.restype = ctypes.c_char_p
    rc = lib.uc2_cast_f32_bf16(None, None, ctypes.c_longlong(0), None)
    assert rc < 0 and lib.uc2_last_error()
    from uc2_b200 import itm
    m = itm.VLXLMRForImageTextRetrieval(cases.config(1), 2048)
    with pytest.raises(RuntimeError):
        m(cases.batch_rank(), compute_loss=False)


def test_state_dict_contract_matches_reference_names():
    """Parameter / state_dict names are the reference's: the shapes table the golden generator
    loaded into the REFERENCE modules is exactly what our modules expose."""
    from uc2_b200 import itm, model
    from uc2_b200.config import pretraining_shapes, retrieval_shapes
    cfg = cases.config(2)
    for fam, Pre, Ret, enc in (("vlxlmr", model.VLXLMRForPretraining, itm.VLXLMRForImageTextRetrieval, "roberta"),
                               ("uniter", model.UniterForPretraining, itm.UniterForImageTextRetrieval, "bert")):
        c = cases.config(2, family=fam)
        m = Pre(c, 2048, 1601)
        names = {n: tuple(p.shape) for n, p in m.named_parameters()}
        assert names == {k: tuple(v) for k, v in pretraining_shapes(c, fam).items()}
        sd = m.state_dict()
        tied = ["feat_regress.weight"] + (["cls.decoder.weight", "cls.decoder.bias"] if fam == "vlxlmr"
                                          else ["cls.predictions.decoder.weight"])
        for t in tied:
            assert t in sd
        assert sd["feat_regress.weight"].data_ptr() == sd[f"{enc}.img_embeddings.img_linear.weight"].data_ptr()
        r = Ret(c, 2048)
        assert {n: tuple(p.shape) for n, p in r.named_parameters()} == \
            {k: tuple(v) for k, v in retrieval_shapes(c, fam).items()}
    # no-decay grouping is name based (optim/misc.py:11): img_layer_norm.weight IS decayed
    from uc2_b200.optim import build_optimizer

    class Opts:
        weight_decay, optim, learning_rate, betas = 0.01, "adamw", 1e-4, (0.9, 0.98)
    m = itm.VLXLMRForImageTextRetrieval(cfg, 2048)
    opt = build_optimizer(m, Opts)
    decayed = {id(p) for p in opt.param_groups[0]["params"]}
    byname = dict(m.named_parameters())
    assert id(byname["roberta.img_embeddings.img_layer_norm.weight"]) in decayed
    assert id(byname["roberta.embeddings.LayerNorm.weight"]) not in decayed
    assert id(byname["roberta.encoder.layer.0.output.dense.bias"]) not in decayed


def test_init_output_and_from_pretrained(tmp_path):
    from uc2_b200 import itm
    import json
    cfg = cases.config(1)
    p = tmp_path / "cfg.json"
    p.write_text(json.dumps(cfg.to_dict()))
    sd = cases.weights(cfg, "retrieval")
    sd["roberta.embeddings.LayerNorm.gamma"] = sd.pop("roberta.embeddings.LayerNorm.weight")   # TF-style names
    sd["roberta.embeddings.LayerNorm.beta"] = sd.pop("roberta.embeddings.LayerNorm.bias")
    m = itm.VLXLMRForImageTextRetrieval.from_pretrained(str(p), sd, img_dim=2048, margin=0.2)
    assert torch.equal(m.roberta.embeddings.LayerNorm.weight, sd["roberta.embeddings.LayerNorm.gamma"])
    m.init_output()
    assert torch.equal(m.rank_output.weight, m.itm_output.weight[1:])
    assert torch.equal(m.rank_output.bias, m.itm_output.bias[1:])


def test_arena_layout_glues_qkv():
    from uc2_b200.arena import _order
    from uc2_b200 import itm
    m = itm.VLXLMRForImageTextRetrieval(cases.config(2), 2048)
    names = _order([n for n, _ in m.named_parameters()])
    i = names.index("roberta.encoder.layer.1.attention.self.query.weight")
    assert [n.split("self.")[1] for n in names[i:i + 6]] == ["query.weight", "key.weight", "value.weight",
                                                            "query.bias", "key.bias", "value.bias"]
    assert sorted(names) == sorted(n for n, _ in m.named_parameters())


def test_lr_schedule_and_loss_reduction():
    from uc2_b200.optim import warmup_linear
    from uc2_b200.train import reduce_loss
    assert warmup_linear(5, 10, 100) == 0.5 and warmup_linear(55, 10, 100) == 0.5
    itm = torch.tensor([1.0, 3.0])
    pos, neg = torch.tensor([2.0]), torch.tensor([1.0, 1.0])
    got = reduce_loss((itm, (pos, neg)), "itm", 0.1)
    assert abs(float(got) - (2.0 + 0.1 * (2.0 - 2.0) / 3)) < 1e-6
    assert float(reduce_loss(torch.tensor([1.0, 2.0]), "mlm")) == 1.5


DP_SCRIPT = r"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, %r)
from uc2_b200 import distributed as D
D.init("gloo")
r, w = D.rank(), D.size()
assert w == 2
flat = torch.arange(1000, dtype=torch.float32) * (r + 1)
sync = D.GradSync(flat, bucket_bytes=256 * 4)
sync.ready(600, 1000)        # "heads"
sync.ready(200, 600)         # "layers"
sync.finish()                # embeddings [0,200) were never reported: finish() must still reduce them
exp = torch.arange(1000, dtype=torch.float32) * 1.5
assert torch.allclose(flat, exp), (flat[:5], exp[:5])
# disabled (gradient accumulation micro-step): nothing may be communicated
flat2 = torch.ones(10) * (r + 1)
s2 = D.GradSync(flat2); s2.enabled = False; s2.ready(0, 10); s2.finish()
assert torch.equal(flat2, torch.ones(10) * (r + 1))
# row-sparse exchange of a [12, 4] table at offset 8 of a flat buffer: rank r touched rows {1, 3+r, 3+r (dup), 7}
flat3 = torch.zeros(8 + 12 * 4 + 5)
flat3[:8] = r + 1.0
tab = flat3[8:8 + 48].view(12, 4)
ids = torch.tensor([7, 1, 3 + r, 3 + r] + [1] * r)      # ragged: rank 1 lists one more (duplicate) id than rank 0
for i in set(ids.tolist()):
    tab[i] = (r + 1) * (i + 1)
flat3[-5:] = 10.0 * (r + 1)
s3 = D.GradSync(flat3, bucket_bytes=16 * 4)
s3.sparse_rows_table(8, 12, 4, ids, pad_row=0)                 # row 0 never carries gradient (padding_idx)
s3.finish()                                   # the rest of the buffer goes through the dense path
exp3 = torch.zeros(12, 4)
exp3[1] = (1 * 2 + 2 * 2) / 2.0; exp3[7] = (1 * 8 + 2 * 8) / 2.0; exp3[3] = 1 * 4 / 2.0; exp3[4] = 2 * 5 / 2.0
assert torch.allclose(flat3[8:56].view(12, 4), exp3), flat3[8:56].view(12, 4)
assert torch.allclose(flat3[:8], torch.full((8,), 1.5)) and torch.allclose(flat3[-5:], torch.full((5,), 15.0))
t = [torch.ones(3) * (r + 1), torch.ones(2) * (r + 1)]
D.all_reduce_and_rescale_tensors(t, 2.0)
assert torch.allclose(t[0], torch.ones(3) * 0.75)
assert D.all_gather_list({"rank": r}) == [{"rank": 0}, {"rank": 1}]
assert D.any_broadcast("task_%%d" %% r, 1) == "task_1"
rows = torch.full((r + 1, 4), float(r))
allr = D.allgather_rows(rows)
assert allr.shape == (3, 4) and allr[0, 0] == 0 and allr[2, 0] == 1
b = torch.zeros(5) + r
D.broadcast_tensors([b], 0)
assert torch.equal(b, torch.zeros(5))
dist.destroy_process_group()
print("ok", r)
"""


def test_data_parallel_logic_gloo_world2(tmp_path):
    script = tmp_path / "dp.py"
    script.write_text(DP_SCRIPT % ROOT)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29731", str(script)],
                         capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-3000:]
    assert out.stdout.count("ok") == 2


def test_recall_eval_matches_reference(golden):
    """uc2_b200.retrieval.itm_eval (vectorised) against eval/itm.py run on the same matrices (tests/golden/retrieval.npz)."""
    from uc2_b200.retrieval import itm_eval
    g = golden("retrieval")
    n_img, cpi = 15, 4
    tids = [f"t{i}" for i in range(n_img * cpi)]
    iids = [f"i{j}" for j in range(n_img)]
    t2i = {t: f"i{i // cpi}" for i, t in enumerate(tids)}
    i2t = {f"i{j}": [tids[j * cpi + k] for k in range(cpi)] for j in range(n_img)}
    log = itm_eval(torch.from_numpy(g["recall|matrix"]), tids, iids, t2i, i2t)
    assert sorted(log) == list(g["recall|keys"])
    np.testing.assert_allclose([log[k] for k in sorted(log)], g["recall|values"], rtol=0, atol=1e-12)
    # and on the score matrix the reference model produced for the small retrieval case
    images, captions, txt_ids, txt2img, img2txts = cases.retrieval_case()
    img_ids = [images[i]["id"] for i in g["retrieval|img_order"]]
    log = itm_eval(torch.from_numpy(g["retrieval|scores"]), txt_ids, img_ids, txt2img, img2txts)
    np.testing.assert_allclose([log[k] for k in sorted(log)], g["retrieval|recall"], rtol=0, atol=1e-12)


def test_dropout_hash_host_mirror():
    """uc2_b200.dropout: the scalar and the numpy mixers agree, keys differ per site / layer / step, the keep rate
    matches p, and thresh / scale describe the same effective probability."""
    from uc2_b200 import dropout as DO
    xs = np.array([0, 1, 12345, 0xFFFFFFFF, 0x9E3779B9], dtype=np.uint32)
    assert [DO.lowbias32(int(x)) for x in xs] == [int(v) for v in DO.lowbias32_np(xs)]
    keys = {DO.site_key(7, c, l, s) for c in (1, 2) for l in (0, 1, 11, 255) for s in range(4)}
    assert len(keys) == 2 * 4 * 4
    for p in (0.1, 0.25, 0.5):
        t = DO.thresh_of(p)
        keep = DO.keep_mask_np(DO.site_key(3, 1, 0, DO.SITE_OUT1), 1 << 20, t)
        assert abs(keep.mean() - (1 - p)) < 2e-3
        assert abs(DO.scale_of(p) * (1 - t / 65536.0) - 1.0) < 1e-12
    assert DO.thresh_of(0.0) == 0 and DO.scale_of(0.0) == 1.0
    with pytest.raises(ValueError):
        DO.thresh_of(1.0)
    a = DO.keep_mask_np(DO.head_key(99, 5), 4096, DO.thresh_of(0.1))
    b = DO.keep_mask_np(DO.head_key(99, 6), 4096, DO.thresh_of(0.1))
    assert (a != b).any()


def test_token_bucket_sampler_and_meta_loader_match_reference(golden):
    """uc2_b200.loader against batches / task sequences produced by the reference's data/sampler.py and
    data/loader.py under the same `random` seeds (tests/golden/loader.npz): index lists bit-exact."""
    import random
    from uc2_b200.loader import MetaLoader, TokenBucketSampler
    g = golden("loader")
    lens = [int(x) for x in g["lens"]]
    for k in range(3):
        bucket, budget, drop = (int(x) for x in g[f"sampler{k}|cfg"])
        random.seed(100 + k)
        batches = [b for b in iter(TokenBucketSampler(lens, bucket, budget, droplast=bool(drop)))]
        assert [len(b) for b in batches] == list(g[f"sampler{k}|sizes"])
        assert [i for b in batches for i in b] == list(g[f"sampler{k}|flat"])
        for b in batches:                      # the token budget holds and full batches are multiples of 8
            assert max(lens[i] for i in b) * len(b) <= budget
        if drop:
            assert all(len(b) % 8 == 0 for b in batches)
    random.seed(7)
    ml = MetaLoader({"mlm": ([1, 2, 3], 2), "itm": [10, 20], "mrfr": ([5], 1)}, accum_steps=3, distributed=False)
    seq = []
    for i, tb in enumerate(ml):
        seq.append(tb)
        if i == 59:
            break
    assert [t for t, _ in seq] == list(g["meta|tasks"])
    assert [b for _, b in seq] == list(g["meta|batches"])
    # distributed mode: every rank draws the same task sequence from its own identically seeded stream
    a = MetaLoader({"x": ([1], 3), "y": [2]}, accum_steps=2, distributed=True, task_seed=11)
    b = MetaLoader({"x": ([1], 3), "y": [2]}, accum_steps=2, distributed=True, task_seed=11)
    ia, ib = iter(a), iter(b)
    assert [next(ia)[0] for _ in range(40)] == [next(ib)[0] for _ in range(40)]
    with pytest.raises(ValueError):
        len(TokenBucketSampler(lens, 8, 100))


class _ScoreStub(torch.nn.Module):
    """Stands in for the retrieval model: returns the scores the batch carries (the bookkeeping is under test)."""
    def forward(self, batch, compute_loss=False):
        return batch["scores"].unsqueeze(1)


def _hn_loader(n_txt, n_img, per_batch, seed, rank=0, world=1):
    g = torch.Generator().manual_seed(seed)
    out = []
    for t in range(rank, n_txt, world):
        gt = torch.Generator().manual_seed(seed * 1000 + t)
        imgs = [f"img{i}" for i in torch.randperm(n_img, generator=gt)[:per_batch].tolist()]
        out.append({"gt_txt_id": f"txt{t}", "neg_img_ids": imgs, "scores": torch.randn(per_batch, generator=gt)})
    return out


def _ref_hard_negs(batches, k):
    """itm.py:385-445 restated literally (single process sees every batch)."""
    from collections import defaultdict
    txt2hardimgs, img_to_score_txts = {}, defaultdict(list)
    for b in batches:
        scores, txt, imgs = b["scores"], b["gt_txt_id"], b["neg_img_ids"]
        txt2hardimgs[txt] = [imgs[i] for i in scores.topk(k, sorted=False)[1].tolist()]
        for i, img in enumerate(imgs):
            img_to_score_txts[img].append((scores[i].item(), txt))
    img2hardtxts = {}
    for img, st in img_to_score_txts.items():
        sc, txts = [s for s, _ in st], [t for _, t in st]
        idx = range(len(txts)) if len(txts) < k else torch.tensor(sc).topk(k, sorted=False)[1].tolist()
        img2hardtxts[img] = [txts[i] for i in idx]
    return txt2hardimgs, img2hardtxts


def test_hard_negative_extraction_and_validate():
    from uc2_b200.retrieval import get_hard_negs, validate
    batches = _hn_loader(40, 30, 12, seed=3)
    t2i, i2t = get_hard_negs(_ScoreStub(), batches, hard_negative_num=5)
    rt2i, ri2t = _ref_hard_negs(batches, 5)
    assert {k: set(v) for k, v in t2i.items()} == {k: set(v) for k, v in rt2i.items()}
    assert {k: set(v) for k, v in i2t.items()} == {k: set(v) for k, v in ri2t.items()}
    # validate (itm.py:447-488): ground truth is index 0 of every batch
    vb = []
    for r, n in ((0, 12), (3, 12), (7, 12), (11, 12), (2, 6)):      # rank of the ground truth inside the batch
        s = torch.arange(n, 0, -1).float()                            # descending: index i has rank i
        s[0], s[r] = s[r].item(), s[0].item()
        vb.append({"scores": s})
    log = validate(_ScoreStub(), vb)
    assert log["valid/recall_1"] == 1 / 5 and log["valid/recall_5"] == 3 / 5 and log["valid/recall_10"] == 4 / 5


HN_SCRIPT = r"""
import sys, torch, torch.distributed as dist
sys.path.insert(0, %r); sys.path.insert(0, %r)
from uc2_b200 import distributed as D
from uc2_b200.retrieval import get_hard_negs
import test_host_cpu as T
D.init("gloo")
r, w = D.rank(), D.size()
mine = T._hn_loader(24, 20, 8, seed=5, rank=r, world=w)
t2i, i2t = get_hard_negs(T._ScoreStub(), mine, hard_negative_num=4)
everything = T._hn_loader(24, 20, 8, seed=5)
rt2i, ri2t = T._ref_hard_negs(everything, 4)
assert {k: set(v) for k, v in t2i.items()} == {k: set(rt2i[k]) for k in t2i} and len(t2i) == 12
if r == 0:
    assert {k: set(v) for k, v in i2t.items()} == {k: set(v) for k, v in ri2t.items()}
else:
    assert i2t == {}
dist.destroy_process_group()
print("ok", r)
"""


def test_hard_negative_extraction_gloo_world2(tmp_path):
    script = tmp_path / "hn.py"
    script.write_text(HN_SCRIPT % (ROOT, os.path.join(ROOT, "tests")))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29741", str(script)],
                         capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-3000:]
    assert out.stdout.count("ok") == 2


class _OutStub(torch.nn.Module):
    """Returns the output the batch carries: the validation bookkeeping is under test, not the model."""
    def forward(self, batch, task=None, compute_loss=True):
        return batch["_out"]


def test_validation_loops_match_reference_formulas():
    """uc2_b200.validate against the formulas of pretrain.py:688-1050 written out with the reference's per-batch
    .item() accumulation."""
    import torch.nn.functional as F
    from uc2_b200 import validate as V
    g = torch.Generator().manual_seed(1)
    m = _OutStub()

    def rn(*s):
        return torch.randn(*s, generator=g)

    # token tasks (validate_mlm / mmxlm / vmlm)
    bs = []
    for n in (7, 12):
        lab = torch.full((4, 9), -1, dtype=torch.long)
        flat = torch.randperm(36, generator=g)[:n]
        lab.view(-1)[flat] = torch.randint(0, 50, (n,), generator=g)
        bs.append({"_out": rn(n, 50), "txt_labels": lab})
    loss = correct = words = 0
    for b in bs:
        l = b["txt_labels"][b["txt_labels"] != -1]
        loss += F.cross_entropy(b["_out"], l, reduction="sum").item()
        correct += (b["_out"].max(dim=-1)[1] == l).sum().item()
        words += l.numel()
    for fn in (V.validate_mlm, V.validate_mmxlm, V.validate_vmlm):
        log = fn(m, bs)
        assert set(log) == {"loss", "acc", "tok_per_s"}
        np.testing.assert_allclose([log["loss"], log["acc"]], [loss / words, correct / words], rtol=1e-6)
    # soft-label tasks and MRC (kl and plain)
    bs = []
    for n in (5, 8):
        tgt = torch.softmax(rn(n, 30) * 3, -1)
        mask = torch.zeros(3, 10, dtype=torch.bool)
        mask.view(-1)[:n] = True
        bs.append({"_out": rn(n, 30), "label_targets": tgt, "tgt_masks": mask, "img_mask_tgt": mask})
    loss = score = feats = 0
    for b in bs:
        p = F.log_softmax(b["_out"], dim=-1)
        loss += F.kl_div(p, b["label_targets"], reduction="sum").item()
        score += (p.max(dim=-1)[1] == b["label_targets"].max(dim=-1)[1]).sum().item()
        feats += b["tgt_masks"].sum().item()
    for log in (V.validate_mmxlm_soft(m, bs), V.validate_vmlm_soft(m, bs), V.validate_mrc(m, bs, "mrc-kl")):
        assert set(log) == {"loss", "acc", "feat_per_s"}
        np.testing.assert_allclose([log["loss"], log["acc"]], [loss / feats, score / feats], rtol=1e-6)
    loss = score = 0
    for b in bs:
        cls = b["label_targets"][:, 1:].max(dim=-1)[1] + 1
        loss += F.cross_entropy(b["_out"], cls, ignore_index=0, reduction="sum").item()
        score += (b["_out"][:, 1:].max(dim=-1)[1] == b["label_targets"][:, 1:].max(dim=-1)[1]).sum().item()
    log = V.validate_mrc(m, bs, "mrc")
    np.testing.assert_allclose([log["loss"], log["acc"]], [loss / feats, score / feats], rtol=1e-6)
    # MRFR
    bs = [{"_out": rn(n, 2048) ** 2, "img_mask_tgt": torch.ones(n, dtype=torch.bool)} for n in (3, 6)]
    log = V.validate_mrfr(m, bs)
    want = sum(b["_out"].sum().item() / 2048 for b in bs) / 9
    np.testing.assert_allclose(log["loss"], want, rtol=1e-6)
    # ITM with the (pos, neg) OT pair
    bs = [{"_out": (rn(n, 2), (rn(3).abs(), rn(n - 3).abs())), "targets": torch.randint(0, 2, (n,), generator=g)}
          for n in (6, 9)]
    loss = score = ot = otp = otn = 0
    for b in bs:
        sc, (p, q) = b["_out"]
        loss += F.cross_entropy(sc, b["targets"], reduction="sum").item()
        score += (sc.max(dim=-1)[1] == b["targets"]).sum().item()
        otp += p.sum().item(); otn += q.sum().item(); ot += p.sum().item() - q.sum().item()
    log = V.validate_itm(m, bs)
    np.testing.assert_allclose([log["valid/loss"], log["valid/acc"], log["valid/ot_loss"], log["valid/ot_pos"],
                                log["valid/ot_neg"]], [loss / 15, score / 15, ot / 15, otp / 15, otn / 15], rtol=1e-5)
    assert "valid/ot_loss" not in V.validate_itm(m, [{"_out": (rn(4, 2), None), "targets": torch.zeros(4, dtype=torch.long)}])
    # dispatcher: prefixes and key naming of pretrain.py:658-685
    out = V.validate(m, {"mrfr_coco": [{"_out": rn(3, 2048), "img_mask_tgt": torch.ones(3, dtype=torch.bool)}]})
    assert set(out["mrfr_coco"]) == {"mrfr_coco_loss", "mrfr_coco_feat_per_s"}


def test_pretrain_loop_bookkeeping_matches_reference_loop():
    """uc2_b200.pretrain_loop.PretrainLoop with the training step stubbed out: meters, counters, logging points,
    validation / checkpoint cadence against pretrain.py:484-656 restated with its per-step .item() calls."""
    import math
    from types import SimpleNamespace
    from uc2_b200.pretrain_loop import PretrainLoop, RunningMeter
    g = torch.Generator().manual_seed(4)
    opts = SimpleNamespace(gradient_accumulation_steps=2, num_train_steps=5, valid_steps=2, grad_norm=2.0,
                           itm_ot_lambda=0.1, ot_pos_only=False, learning_rate=1e-4, decay="linear", warmup_steps=2)
    stream = []
    for i in range(14):
        B, S = 4 + i % 3, 10
        attn = (torch.rand(B, S, generator=g) > 0.3).long()
        if i % 3 == 0:
            npos = 0 if i == 6 else 2                                   # one batch without positives: NaN mean dropped
            out = (torch.rand(B, generator=g), (torch.rand(npos, generator=g), torch.rand(B - npos, generator=g)))
            name = "itm_coco"
        else:
            out = torch.rand(3 + i, generator=g)
            name = "mlm_coco" if i % 3 == 1 else "mrfr_vg"
        stream.append((name, {"input_ids": torch.zeros(B, 6, dtype=torch.long), "attn_masks": attn, "_out": out}))

    class FakeStep(object):
        def __init__(self):
            self.global_step, self.micro, self.last_out, self.last_grad_norm = 0, 0, None, None
        def __call__(self, batch, task):
            from uc2_b200.train import reduce_loss
            self.last_out = batch["_out"]
            self.micro += 1
            if self.micro % opts.gradient_accumulation_steps == 0:
                self.global_step += 1
                self.last_grad_norm = torch.tensor(1.5)
            return reduce_loss(batch["_out"], task, opts.itm_ot_lambda)

    class Saver(object):
        calls = []
        def save(self, model, step, optimizer=None):
            self.calls.append((step, optimizer is not None))

    class Restorer(object):
        global_step, n = 0, 0
        def step(self):
            Restorer.n += 1

    logged = []
    model = torch.nn.Linear(1, 1)
    loop = PretrainLoop(model, SimpleNamespace(param_groups=[{"lr": 0.1}]), opts, val_dataloaders={},
                        model_saver=Saver(), restorer=Restorer(), scalar_log=lambda n, v, s: logged.append((n, v, s)),
                        log_every=2, step_fn=FakeStep())
    end = loop.run(iter(stream), task_names=["itm_coco", "mlm_coco", "mrfr_vg"])
    assert end == 5
    # ---- the reference loop, literally
    meters = {}
    def meter(k):
        return meters.setdefault(k, RunningMeter(f"loss/{k}"))
    for t in ("itm_coco", "mlm_coco", "mrfr_vg", "itm_coco_xe", "itm_coco_ot", "itm_coco_ot_pos", "itm_coco_ot_neg"):
        meter(t)
    n_ex, n_in, n_l = {}, {}, {}
    gs, snaps = 0, {}
    for step, (name, b) in enumerate(stream):
        n_ex[name] = n_ex.get(name, 0) + b["input_ids"].size(0)
        n_in[name] = n_in.get(name, 0) + (b["attn_masks"] == 1).sum().item()
        loss = b["_out"]
        if name.startswith("itm"):
            itm_loss, (pos, neg) = loss
            n_l[name] = n_l.get(name, 0) + itm_loss.size(0)
            itm_loss = itm_loss.mean()
            ot = (pos.sum() - neg.sum()) / (pos.size(0) + neg.size(0))
            p = pos.mean().item()
            if not math.isnan(p):
                meter(f"{name}_ot_pos")(p)
            q = neg.mean().item()
            if not math.isnan(q):
                meter(f"{name}_ot_neg")(q)
            loss = itm_loss + opts.itm_ot_lambda * ot
            meter(f"{name}_xe")(itm_loss.item())
            meter(f"{name}_ot")(ot.item())
        else:
            n_l[name] = n_l.get(name, 0) + loss.size(0)
            loss = loss.mean()
        meter(name)(loss.item())
        if (step + 1) % 2 == 0:
            gs += 1
            snaps[gs] = {m.name: m.val for m in meters.values() if m.val is not None}
        if gs >= opts.num_train_steps:
            break
    assert dict(loop.n_examples) == n_ex and dict(loop.n_in_units) == n_in and dict(loop.n_loss_units) == n_l
    for k, m in meters.items():
        got = loop.task2loss[k].val
        assert (got is None) == (m.val is None), k
        if got is not None:
            np.testing.assert_allclose(got, m.val, rtol=1e-6, err_msg=k)
    # logging points: every 2 optimizer steps the meters as of that step
    for s in (2, 4):
        got = {n: v for n, v, st in logged if st == s and n.startswith("loss/")}
        assert set(got) == set(snaps[s])
        for n in got:
            np.testing.assert_allclose(got[n], snaps[s][n], rtol=1e-6, err_msg=f"{n}@{s}")
        assert ("grad_norm", 1.5, s) in logged
        assert {n for n, _, st in logged if st == s and n.startswith("perf/")} == {
            f"perf/{t}_{u}_per_s" for t in ("itm_coco", "mlm_coco", "mrfr_vg") for u in ("ex", "in", "loss")}
    assert [s for n, _, s in logged if n == "lr"] == [1, 2, 3, 4, 5]
    # validation + checkpoint at steps 2 and 4 (with optimizer state) and once more at the end (5, model only)
    assert Saver.calls == [(2, True), (4, True), (5, False)] and Restorer.n == 5
    assert model.training


def test_lr_schedules_match_reference(golden):
    """optim/sched.py (tests/golden/sched.npz was produced by the reference module itself)."""
    from types import SimpleNamespace
    from uc2_b200.optim import get_lr_sched, get_xlmr_lr_sched
    g = golden("sched")
    for decay in ("linear", "invsqrt", "constant", "vqa"):
        o = SimpleNamespace(decay=decay, learning_rate=3e-4, xlmr_lr=1e-5, warmup_steps=10, num_train_steps=50,
                            warm_int=4, decay_int=7, decay_st=20, decay_rate=0.5)
        np.testing.assert_array_equal([get_lr_sched(s, o) for s in range(60)], g[f"{decay}|lr"])
        np.testing.assert_array_equal([get_xlmr_lr_sched(s, o) for s in range(60)], g[f"{decay}|xlmr_lr"])


def test_finetune_loop_control_flow():
    """uc2_b200.pretrain_loop.FinetuneLoop against itm.py:253-358: separate learning rates, validation / checkpoint
    cadence, loader rebuild after hard-negative mining, accumulation."""
    from types import SimpleNamespace
    from uc2_b200.optim import get_lr_sched, get_xlmr_lr_sched
    from uc2_b200.pretrain_loop import FinetuneLoop, RunningMeter
    opts = SimpleNamespace(gradient_accumulation_steps=2, num_train_steps=7, valid_steps=3, grad_norm=2.0,
                           learning_rate=1e-3, xlmr_lr=1e-5, decay="linear", warmup_steps=2, separate_lr=True,
                           steps_per_hard_neg=4)
    optimizer = SimpleNamespace(param_groups=[{"lr": 0.0} for _ in range(4)])
    g = torch.Generator().manual_seed(2)
    losses = torch.rand(64, generator=g)

    class FakeStep(object):
        def __init__(self):
            self.global_step, self.micro, self.last_grad_norm, self.seen = 0, 0, None, []
        def __call__(self, batch, task):
            assert task is None
            self.seen.append(batch["i"])
            self.micro += 1
            if self.micro % 2 == 0:
                self.global_step += 1
                self.last_grad_norm = torch.tensor(0.5)
            return losses[batch["i"]]

    built, events, logged = [], [], []
    def build_loader():
        built.append(len(built))
        base = 20 * (len(built) - 1)
        return [{"input_ids": torch.zeros(3, 5, dtype=torch.long), "i": base + k} for k in range(20)]

    class Saver(object):
        def save(self, model, step, optimizer=None):
            events.append(("save", step))

    step = FakeStep()
    loop = FinetuneLoop(torch.nn.Linear(1, 1), optimizer, opts, build_loader,
                        validate_fn=lambda m: events.append(("val", step.global_step)) or {"valid/recall_1": 0.5},
                        hard_neg_fn=lambda m: events.append(("hn", step.global_step)), model_saver=Saver(),
                        scalar_log=lambda n, v, s: logged.append((n, v, s)), log_every=2, step_fn=step)
    assert loop.run() == 7
    # loader 0 runs until optimizer step 4 (8 micro-steps), then hard negatives are mined and the loader is rebuilt
    assert built == [0, 1] and step.seen == list(range(8)) + list(range(20, 26))
    assert events == [("val", 3), ("save", 3), ("hn", 4), ("val", 6), ("save", 6)]
    assert loop.n_examples == 14 * 3
    ref = RunningMeter("loss")
    for i in step.seen:
        ref(losses[i].item())
    np.testing.assert_allclose(loop.running_loss.val, ref.val, rtol=1e-6)
    for s in range(1, 8):
        assert ("lr", get_lr_sched(s, opts), s) in logged and ("xlmr_lr", get_xlmr_lr_sched(s, opts), s) in logged
    assert [s for n, _, s in logged if n == "perf/ex_per_s"] == [2, 4, 6]
    assert ("valid/recall_1", 0.5, 3) in logged
    # the per-group learning rates TrainStep would install
    lrs = loop.lr_fn(3)
    assert lrs == [get_xlmr_lr_sched(3, opts)] * 2 + [get_lr_sched(3, opts)] * 2


def test_sampling_draws_match_reference(golden):
    """uc2_b200.sampling under the seeds of tests/golden/make_golden.py::case_sampling (the reference's own
    random_word / _get_img_mask / sample_negative produced the fixture): same `random` stream, same draws."""
    import random
    from uc2_b200 import sampling as S, synth
    g = golden("sampling")
    lens = [int(x) for x in g["random_word|lens"]]
    random.seed(11)
    toks, labs = [], []
    for k, n in enumerate(lens):
        ids = [int(x) for x in synth.det_randint(n, 5, 250001, 300 + k, 2)]
        t, l = S.random_word(ids, (5, 250001), 250001)
        assert t is ids                                                   # edited in place, like the reference
        toks += t
        labs += l
    np.testing.assert_array_equal(toks, g["random_word|tokens"])
    np.testing.assert_array_equal(labs, g["random_word|labels"])
    assert any(l != -1 for l in labs[:1])                                 # a 1-token sentence still gets its mask
    random.seed(12)
    flat = np.concatenate([S.get_img_mask(0.15, int(n)).numpy().astype(np.uint8) for n in g["img_mask|nbbs"]])
    np.testing.assert_array_equal(flat, g["img_mask|flat"])
    random.seed(13)
    imgs = [f"img{i}" for i in range(12)]
    txts = [f"txt{i}" for i in range(36)]
    img2txts = {f"img{i}": [f"txt{3 * i + k}" for k in range(3)] for i in range(12)}
    pairs = []
    for t in (0, 7, 20, 35, 14):
        gi = f"img{t // 3}"
        for ns in (1, 2):
            pairs += [f"{a}|{b}" for a, b in S.rank_id_pairs(f"txt{t}", gi, imgs, txts, img2txts[gi], ns)]
    assert pairs == list(g["rank|pairs"])
    ids, lab = S.create_mlm_io([7, 8, 9], (5, 100), 99, 0, 2)
    assert ids[0] == 0 and ids[-1] == 2 and lab[0] == -1 and lab[-1] == -1 and ids.numel() == lab.numel() == 5


def test_in_memory_datasets_host_logic():
    """uc2_b200.datasets: sample construction over a (stubbed) feature arena -- lengths for the token-bucket sampler,
    rank sharding, masking alignment, negative sampling invariants, collate dispatch."""
    import random
    from uc2_b200 import datasets as DS
    from uc2_b200.loader import TokenBucketSampler
    n_img, cpi = 9, 3
    class Arena(object):                       # the only two things the datasets ask of a FeatureArena
        nbb = [10 + 7 * i for i in range(n_img)]
        def __len__(self):
            return n_img
    arena = Arena()
    names = [f"img{i}" for i in range(n_img)]
    ex = {f"t{k}": {"input_ids": [5 + (k * 13 + j) % 200 for j in range(3 + k % 9)], "img_fname": f"img{k // cpi}"}
          for k in range(n_img * cpi)}
    ex["too_long"] = {"input_ids": list(range(5, 80)), "img_fname": "img0"}
    db = DS.TextDB(ex, max_txt_len=60, mask=999, v_range=(5, 900))
    assert "too_long" not in db.ids and len(db.ids) == 27
    assert db.combine_inputs([7, 8]).tolist() == [0, 7, 8, 2] and db.img2txts["img2"] == ["t6", "t7", "t8"]
    idx = DS.ImageIndex(arena, names)
    base = DS.MrfrDataset(0.15, db, idx, rank=1, world=2)
    assert base.ids == db.ids[1::2]
    assert base.lens == [len(ex[i]["input_ids"]) + idx.name2nbb[ex[i]["img_fname"]] for i in base.ids]
    random.seed(0)
    ids, k, m = base[3]
    assert ids[0] == 0 and ids[-1] == 2 and m.dtype == torch.bool and m.numel() == arena.nbb[k] and bool(m.any())
    mlm = DS.MlmDataset(db, idx)
    a, lab, k = mlm[4]
    assert a.numel() == lab.numel() == len(ex["t4"]["input_ids"]) + 2 and lab[0] == -1 and lab[-1] == -1
    changed = (a[1:-1] != torch.tensor(ex["t4"]["input_ids"])).nonzero().squeeze(1) + 1
    assert bool((lab[changed] != -1).all()) and bool((lab != -1).any()) and k == 1
    assert ex["t4"]["input_ids"] == [5 + (4 * 13 + j) % 200 for j in range(7)]      # the store itself is untouched
    # ITM: labels, negatives and lens move together and are reproducible under the seeds
    np.random.seed(3); random.seed(3)
    itm = DS.ItmDataset(db, idx, neg_sample_p=0.5)
    for i, id_ in enumerate(itm.ids):
        own = ex[id_]["img_fname"]
        assert (itm.train_imgs[i] == own) == (itm.labels[i] == 1)
        assert itm.lens[i] == len(ex[id_]["input_ids"]) + idx.name2nbb[itm.train_imgs[i]]
    snap = (list(itm.labels), list(itm.train_imgs))
    np.random.seed(3); random.seed(3)
    itm.new_epoch()
    assert (list(itm.labels), list(itm.train_imgs)) == snap and 0 < sum(snap[0]) < 27
    batches = list(iter(TokenBucketSampler(itm.lens, 8, 1000, droplast=False)))
    assert sorted(i for b in batches for i in b) == list(range(27))
    # rank items: positive first, wrong images, then captions of other images
    rk = DS.ItmRankDataset(db, idx, neg_sample_size=2)
    item = rk[5]
    assert len(item) == 5 and item[0][1] == 1
    assert all(k != 1 for _, k in item[1:3]) and all(k == 1 for _, k in item[3:])
    own_caps = [db.combine_inputs(ex[t]["input_ids"]).tolist() for t in ("t3", "t4", "t5")]
    assert item[0][0].tolist() == own_caps[2] and all(p[0].tolist() not in own_caps for p in item[3:])

    class Rec(object):
        def __getattr__(self, name):
            return lambda *a, **k: (name, a, k)
    name, args, _ = rk.collate(Rec(), [item, rk[6]])
    assert name == "itm_rank" and len(args[0]) == 10 and args[2] == 5
    name, args, kw = DS.ItmDataset.collate(Rec(), [itm[0], itm[1]], with_ot=False)
    assert name == "itm" and args[2] == [int(itm.labels[0]), int(itm.labels[1])] and kw == {"with_ot": False}
    assert DS.MrcDataset.collate(Rec(), [base[0]])[0] == "mrc" and DS.MlmDataset.collate(Rec(), [mlm[0]])[0] == "mlm"


def test_checkpoint_dict_helpers():
    from uc2_b200.save import inject_early_adaptation, rename_checkpoint
    ck = {"embeddings.word_embeddings.weight": 1, "encoder.layer.0.output.dense.bias": 2}
    assert rename_checkpoint(ck) is ck and sorted(ck) == ["bert.embeddings.word_embeddings.weight",
                                                          "bert.encoder.layer.0.output.dense.bias"]
    w, b = torch.ones(768, 2048), torch.zeros(768)
    out = inject_early_adaptation({}, {"v2w_linear.weight": w, "v2w_linear.bias": b})
    assert out["roberta.img_embeddings.img_linear.weight"] is w and out["roberta.img_embeddings.img_linear.bias"] is b


def test_fp16_compressed_checkpoint_loads():
    """utils/save.py:147-162 stores restore.pt with every float tensor in fp16: such a state dict must load into the
    fp32 parameters (the optimizer moments are cast the same way on the device)."""
    from uc2_b200 import itm
    from uc2_b200.save import _to_cpu
    cfg = cases.config(1)
    sd = cases.weights(cfg, "retrieval")
    packed = _to_cpu({"a": sd, "n": 3, "l": [torch.ones(2), torch.arange(3)]}, half=True)
    assert packed["a"]["rank_output.weight"].dtype == torch.float16 and packed["l"][1].dtype == torch.int64
    assert packed["n"] == 3 and _to_cpu(sd)["rank_output.weight"].dtype == torch.float32
    m = itm.VLXLMRForImageTextRetrieval(cfg, 2048)
    m.load_state_dict(packed["a"], strict=False)
    p = dict(m.named_parameters())["rank_output.weight"]
    assert p.dtype == torch.float32
    torch.testing.assert_close(p.detach(), sd["rank_output.weight"].half().float())


def _digest(t):
    f = t.detach().double()
    flat = f.reshape(-1)
    idx = torch.linspace(0, flat.numel() - 1, 64).long()
    return np.concatenate([np.array(f.shape, dtype=np.float64), f.reshape(f.size(0), -1).sum(1).numpy(), flat[idx].numpy()])


def _check_batch(g, tag, batch, skip=()):
    keys = [k for k in g.files if k.startswith(tag + "|")]
    assert keys
    for k in keys:
        path = k.split("|")[1:]
        if path[0] in skip:
            continue
        v = batch
        for p in path:
            v = v[p]
        if torch.is_tensor(v):
            if v.is_floating_point():
                np.testing.assert_allclose(_digest(v), g[k], rtol=0, atol=0, err_msg=k)
            else:
                np.testing.assert_array_equal(v.numpy().astype(np.int64), g[k], err_msg=k)
        else:
            assert v == g[k][0], k
    ours = set(batch) - {"n_masked"}
    theirs = {k.split("|")[1] for k in keys}
    assert ours == theirs, (tag, ours ^ theirs)


def test_whole_collates_match_reference(golden):
    """uc2_b200.batch.collate_* against the reference's complete collate functions (tests/golden/collate.npz was
    written by data/itm.py, data/mrm.py and data/mlm.py themselves): every key, integer tensors bit-exact, float
    tensors by digest (shape, per-sample sums, strided samples), masks as 0/1."""
    from uc2_b200 import batch as B, synth
    g = golden("collate")
    items = cases._items(6, 71, cases.SMALL_VOCAB, "vlxlmr")
    targets = [1, 0, 0, 1, 1, 0]
    _check_batch(g, "itm_ot", B.collate_itm(items, targets, with_ot=True))
    _check_batch(g, "itm", B.collate_itm(items, targets, with_ot=False))
    _check_batch(g, "rank", B.collate_itm_rank(items, 6))
    lab = synth.make_mlm_labels([it["input_ids"] for it in items], 71, mask_id=cases.SMALL_VOCAB - 1,
                                vocab=cases.SMALL_VOCAB)
    _check_batch(g, "mlm", B.collate_mlm(items, lab))
    nbbs = [it["img_feat"].size(0) for it in items]
    masks = synth.make_img_masks(nbbs, 71)
    _check_batch(g, "mrfr", B.collate_mrfr(items, masks))
    soft = [synth.make_soft_labels(nb, 71 * 31 + i) for i, nb in enumerate(nbbs)]
    _check_batch(g, "mrc", B.collate_mrc(items, masks, soft))
    _check_batch(g, "tlm", B.collate_tlm(items, lab))
    img_lab = []
    for i, (nb, mk) in enumerate(zip(nbbs, masks)):
        tok = torch.from_numpy(synth.det_randint(nb, 5, cases.SMALL_VOCAB, 71 * 77 + i, 5).astype(np.int64))
        img_lab.append(torch.where(mk, tok, torch.full_like(tok, -1)))
    _check_batch(g, "mmxlm", B.collate_mmxlm(items, lab, masks, img_lab))
    tok_soft = [torch.softmax(torch.from_numpy(synth.det_normal((nb, 16), 71 * 5 + i, 2.0)).float(), -1)
                for i, nb in enumerate(nbbs)]
    _check_batch(g, "mmxlm_soft", B.collate_mmxlm_soft(items, masks, tok_soft))


def test_attention_dropout_block_stream_host_mirror():
    """uc2_b200/dropout.py attn_keep_mask_np (the tcgen05 attention kernels' mask, csrc/common.cuh): deterministic, the
    documented formula element by element, keep rate p, and no structure a dropout user would notice (block drop counts
    binomial, neighbours uncorrelated)."""
    from uc2_b200 import dropout as DO
    p, S = 0.1, 160
    t = DO.thresh_of(p)
    key = DO.head_key(0xABCDEF, 7)
    m = DO.attn_keep_mask_np(key, S, t)
    assert m.shape == (S, S) and np.array_equal(m, DO.attn_keep_mask_np(key, S, t))
    for i, j in ((0, 0), (5, 9), (17, 31), (159, 159), (100, 3)):
        h = DO.lowbias32(key ^ (((i >> 4) << 16) | (j >> 4)))
        e = (h * pow(DO.ATTN_CA, i & 15, 1 << 32) * pow(DO.ATTN_CB, j & 15, 1 << 32)) & DO.M32
        assert bool(m[i, j]) == (e >= (t << 16))
    masks = np.stack([DO.attn_keep_mask_np(DO.head_key(0x1234, bh), S, t) for bh in range(96)])
    assert abs(masks.mean() - (1 - DO.thresh_of(p) / 65536.0)) < 2e-3
    drops = (~masks).reshape(96, 10, 16, 10, 16).sum((2, 4)).ravel()
    assert abs(drops.var() / (256 * 0.1 * 0.9) - 1) < 0.1                     # binomial block counts
    flat = masks.astype(np.float64)
    assert abs(np.corrcoef(flat[:, :, :-1].ravel(), flat[:, :, 1:].ravel())[0, 1]) < 5e-3
    assert abs(np.corrcoef(flat[:, :-1].ravel(), flat[:, 1:].ravel())[0, 1]) < 5e-3
    assert DO.attn_keep_mask_np(key, 33, 0).all()                              # thresh 0: dropout off


def test_bench_workloads_host_side():
    """bench.py builds the BASELINE configs it names: cfg3 task cycle at 64 x (60 + 100), cfg5 VTLM pairs at S = 222 with
    TLM position ids restarting at the second <s> (data/mlm.py:420-428), and one row of work per workload."""
    import bench
    assert set(bench.WORKLOADS) == {"pretrain", "itm", "vtlm"}
    hb = bench.host_batches("vtlm", seed=3, n=4)
    assert [t for t, _ in hb] == ["tlm"]
    b = hb[0][1]
    assert b["attn_masks"].shape == (4, 222) and b["input_ids"].shape == (4, 122)
    pos = b["position_ids"][0].tolist()
    assert pos[:3] == [2, 3, 4] and pos[61] == 2 and pos[60] == 62          # restart at the second <s>
    assert b["n_masked"] == int((b["txt_labels"] != -1).sum()) > 0
    assert int(b["txt_labels"][:, 60].max()) == -1 and int(b["txt_labels"][:, 61].max()) == -1
    pb = bench.host_batches("pretrain", seed=3, n=4)
    assert [t for t, _ in pb] == list(bench.TASKS)
    assert all(x["attn_masks"].shape == (4, 160) for _, x in pb)


def test_bench_reference_arm_line_contract():
    """`bench.py --impl reference` (the CPU arm: the oracle port of the reference step on the host cores) prints ONE JSON
    line with the keys the driver reads.  Two layers and one step keep it to well under a minute of host time; the
    workload shape (64 x (60 + 100), vocabulary 250 002) is the real one."""
    import json
    import subprocess
    import sys
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0", "--layers", "2"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, out.stdout[-2000:]
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["metric"].startswith("pretrain samples/sec") and d["unit"] == "samples/s"
    assert d["steps"] == 1 and d["n_gpus"] == 1 and d["value"] > 0
    assert abs(d["value"] - 64 / (d["ms_per_step"] / 1e3)) <= 1e-6 * d["value"]          # 64 samples per step
    assert "workload" in d["config"] and d["config"].get("per_gpu_batch") == 64
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "sample" in cb
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
