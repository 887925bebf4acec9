"""CPU: the oracle restatement (oracle/uc2_oracle.py) reproduces what the reference's own
modules produced (tests/golden/*.npz, made by tests/golden/make_golden.py)."""
import numpy as np
import pytest
import torch

import cases
from oracle import uc2_oracle as O
from uc2_b200 import batch as B


def digest(t):
    f = t.detach().reshape(-1).double()
    idx = torch.linspace(0, f.numel() - 1, 48).long()
    return np.concatenate([[f.norm().item()], f[idx].numpy()])


def grad_sd(sd):
    return {k: v.clone().requires_grad_(True) for k, v in sd.items()}


def check_grads(g, tag, sd, rtol=2e-4):
    keys = [k for k in g.files if k.startswith(f"{tag}|grad|")]
    assert keys
    seen = 0
    for k in keys:
        name = k.split("|")[2]
        if name not in sd:          # tied alias of a tensor we hold under its primary name
            continue
        assert sd[name].grad is not None, name
        ref = g[k]
        got = digest(sd[name].grad)
        scale = max(ref[0] / np.sqrt(sd[name].numel()), 1e-12)
        np.testing.assert_allclose(got[0], ref[0], rtol=rtol, atol=1e-9, err_msg=name)
        np.testing.assert_allclose(got[1:], ref[1:], rtol=0, atol=50 * rtol * scale + 1e-9, err_msg=name)
        seen += 1
    # params the reference left without grad must have none here either
    with_grad = {k.split("|")[2] for k in keys}
    for n, p in sd.items():
        if p.grad is not None and p.grad.abs().sum() > 0:
            assert n in with_grad, f"oracle produced a gradient the reference did not: {n}"
    assert seen > 10


@pytest.mark.parametrize("family", ["vlxlmr", "uniter"])
def test_pretraining_tasks(golden, family):
    g = golden("pretrain" if family == "vlxlmr" else "pretrain_uniter")
    cfg = cases.config(2, family=family)
    fam = O.Family(family)
    sd0 = cases.weights(cfg, "pretrain", family)
    batches = {"itm": cases.batch_itm(family=family), "mlm": cases.batch_mlm(family=family),
               "mrfr": cases.batch_mrfr(family=family), "mrc-kl": cases.batch_mrc(family=family)}
    batches["mrc"] = batches["mrc-kl"]
    if family == "vlxlmr":
        batches["mmxlm"] = cases.batch_mmxlm()
        batches["vmlm-soft"] = cases.batch_mmxlm_soft()
    # packed embedding + hidden states
    for tag in ("itm", "mrfr"):
        b = batches[tag]
        pos = b["position_ids"] if family == "uniter" else None
        with torch.no_grad():
            hs = O.encoder(sd0, fam, b["input_ids"], pos, b["img_feat"], b["img_pos_feat"], b["attn_masks"],
                           b["gather_index"], b.get("img_masks"), all_layers=True)
        np.testing.assert_allclose(hs[0].numpy(), g[f"{tag}|emb"], atol=2e-5)
        rows = g[f"{tag}|hidden_rows"]
        got = np.stack([h[:, rows].numpy() for h in hs[1:]])
        np.testing.assert_allclose(got, g[f"{tag}|hidden"], atol=5e-5)
    for task, b in batches.items():
        sd = grad_sd(sd0)
        out = O.forward_pretraining(sd, fam, b, task)
        if task == "itm":
            np.testing.assert_allclose(out[0].detach().numpy(), g["itm|itm_loss"], atol=1e-5)
            np.testing.assert_allclose(out[1][0].detach().numpy(), g["itm|ot_pos"], rtol=1e-4, atol=1e-6)
            np.testing.assert_allclose(out[1][1].detach().numpy(), g["itm|ot_neg"], rtol=1e-4, atol=1e-6)
        else:
            np.testing.assert_allclose(out.detach().numpy(), g[f"{task}|loss_vec"], rtol=1e-4, atol=1e-5)
        loss = O.pretraining_loss(out, task, itm_ot_lambda=0.1)
        np.testing.assert_allclose(loss.item(), g[f"{task}|loss"][0], rtol=1e-5)
        loss.backward()
        check_grads(g, task, sd)
        with torch.no_grad():
            sc = O.forward_pretraining(sd0, fam, b, task, compute_loss=False)
            sc = sc[0] if task == "itm" else sc
            sc = sc.numpy()
            sc = sc if sc.shape[-1] <= 2048 else sc[:, ::97]
        np.testing.assert_allclose(sc, g[f"{task}|scores"], atol=5e-5)


def test_rank(golden):
    g = golden("rank")
    cfg = cases.config(2)
    fam = O.Family("vlxlmr")
    sd = grad_sd(cases.weights(cfg, "retrieval"))
    b = cases.batch_rank()
    loss = O.forward_retrieval(sd, fam, b)
    np.testing.assert_allclose(loss.detach().numpy(), g["rank|loss_mat"], atol=1e-6)
    loss.mean().backward()
    check_grads(g, "rank", sd)
    with torch.no_grad():
        np.testing.assert_allclose(O.forward_retrieval(sd, fam, b, compute_loss=False).numpy(),
                                   g["rank|scores"], atol=5e-5)


def test_cfg1_full_size(golden):
    """BASELINE.json configs[0] at full size (12 layers, XLM-R vocabulary)."""
    g = golden("cfg1")
    cfg = cases.config(12, vocab=250002)
    fam = O.Family("vlxlmr")
    sd = cases.weights(cfg, "retrieval")
    b = cases.batch_rank(n=8, sample_size=1, seed=42, vocab=250002, txt_len=40, num_bb=36)
    with torch.no_grad():
        hs = O.encoder(sd, fam, b["input_ids"], None, b["img_feat"], b["img_pos_feat"], b["attn_masks"],
                       b["gather_index"], all_layers=True)
        rows = g["cfg1|hidden_rows"]
        np.testing.assert_allclose(hs[0][:, rows].numpy(), g["cfg1|emb"], atol=2e-5)
        np.testing.assert_allclose(np.stack([h[:, rows].numpy() for h in hs[1:]]), g["cfg1|hidden"], atol=2e-4)
        sc = O.forward_retrieval(sd, fam, b, compute_loss=False)
    np.testing.assert_allclose(sc.numpy(), g["cfg1|scores"], atol=1e-4)


def test_index_builders(golden):
    g = golden("index")
    tls, nbs = [int(x) for x in g["lens"][0]], [int(x) for x in g["lens"][1]]
    T, S = max(tls), max(t + n for t, n in zip(tls, nbs))
    # oracle (loops) and product host code (closed form) against the reference's output, bit-exact
    assert np.array_equal(O.gather_index(tls, nbs, T, S), g["gather_index"])
    assert np.array_equal(B.get_gather_index(tls, nbs, len(tls), T, S).numpy(), g["gather_index"])
    assert np.array_equal(O.ot_scatter(tls, T, S), g["ot_scatter"])
    assert np.array_equal(B.compute_ot_scatter(tls, T, S).numpy(), g["ot_scatter"])
    assert np.array_equal(O.pad_mask(tls, T), g["txt_pad"].astype(bool))
    assert np.array_equal(B.compute_pad(nbs, max(nbs)).numpy(), g["img_pad"].astype(bool))
    assert np.array_equal(O.position_ids_from_input_ids(g["pos_in"], 1), g["pos_out"])
    feats = [torch.arange(n * 3, dtype=torch.float32).view(n, 3) + i for i, n in enumerate(nbs)]
    assert np.array_equal(B.pad_tensors(feats, nbs).numpy(), g["pad_tensors"])


def test_ot(golden):
    g = golden("ot")
    tls, nbs = [int(x) for x in g["lens"][0]], [int(x) for x in g["lens"][1]]
    txt = torch.from_numpy(cases.synth.det_normal((4, 12, 768), 901)).requires_grad_(True)
    img = torch.from_numpy(cases.synth.det_normal((4, 20, 768), 902)).requires_grad_(True)
    d = O.optimal_transport_dist(txt, img, B.compute_pad(tls, 12), B.compute_pad(nbs, 20))
    np.testing.assert_allclose(d.detach().numpy(), g["dist"], rtol=1e-5)
    d.sum().backward()
    np.testing.assert_allclose(txt.grad.numpy(), g["dtxt"], atol=1e-7)
    np.testing.assert_allclose(img.grad.numpy(), g["dimg"], atol=1e-7)


def test_adamw(golden):
    g = golden("adamw")
    names = ["enc.dense.weight", "enc.dense.bias", "enc.LayerNorm.weight", "img_layer_norm.weight"]
    shapes = [(33, 17), (17,), (17,), (17,)]
    ps = [torch.from_numpy(cases.synth.det_normal(s, 700 + i, 0.5)) for i, s in enumerate(shapes)]
    ms = [torch.zeros_like(p) for p in ps]
    vs = [torch.zeros_like(p) for p in ps]
    # name-based grouping: img_layer_norm.weight IS decayed (SURVEY section 7 "hard parts")
    assert [O.no_decay(n) for n in names] == [False, True, True, False]
    assert list(g["decay_flags"]) == [2, 2]
    for step in range(1, 5):
        lr = 3e-3 * O.warmup_linear(step, 2, 10)
        gs = [torch.from_numpy(cases.synth.det_normal(p.shape, 800 + 10 * step + i, 2.0)) for i, p in enumerate(ps)]
        gn = O.clip_grad_norm(gs, 5.0)
        np.testing.assert_allclose(float(gn), g[f"gnorm{step}"][0], rtol=1e-6)
        for n, p, gr, m, v in zip(names, ps, gs, ms, vs):
            O.adamw_step(p, gr, m, v, step, lr, 0.9, 0.98, 1e-6, 0.0 if O.no_decay(n) else 0.01)
            np.testing.assert_allclose(p.numpy(), g[f"p{step}|{n}"], rtol=1e-5, atol=1e-7)


def test_dropout_multipliers_reach_every_task_forward():
    """The oracle's `drop=` multipliers (the explicit stand-in for nn.Dropout in training mode) thread through the
    task forwards: all-ones masks reproduce the no-dropout result exactly, real masks change it, and the scaling of a
    kept element is 1 / (1 - p)."""
    cfg = cases.config(1)
    fam = O.Family("vlxlmr")
    sd = cases.weights(cfg, "pretrain")
    b = cases.batch_mrfr()
    B_, T, R, S = b["input_ids"].size(0), b["input_ids"].size(1), b["img_feat"].size(1), b["attn_masks"].size(1)
    ones = {"emb": torch.ones(B_, T + R, 768), "layers": [(torch.ones(B_, 12, S, S), torch.ones(B_, S, 768),
                                                           torch.ones(B_, S, 768))]}
    base = O.forward_pretraining(sd, fam, b, "mrfr")
    same = O.forward_pretraining(sd, fam, b, "mrfr", drop=ones)
    assert torch.equal(base, same)
    g = torch.Generator().manual_seed(0)
    p = 0.1
    mk = lambda *s: torch.bernoulli(torch.full(s, 1 - p), generator=g) / (1 - p)
    real = {"emb": mk(B_, T + R, 768), "layers": [(mk(B_, 12, S, S), mk(B_, S, 768), mk(B_, S, 768))]}
    vals = real["emb"].unique().tolist()
    assert len(vals) == 2 and vals[0] == 0.0 and abs(vals[1] - 1 / (1 - p)) < 1e-6
    diff = O.forward_pretraining(sd, fam, b, "mrfr", drop=real)
    assert not torch.allclose(base, diff)
    rsd = cases.weights(cfg, "retrieval")
    rb = cases.batch_rank()
    S2 = rb["attn_masks"].size(1)
    n = rb["input_ids"].size(0)
    ones2 = {"emb": torch.ones(n, rb["input_ids"].size(1) + rb["img_feat"].size(1), 768),
             "layers": [(torch.ones(n, 12, S2, S2), torch.ones(n, S2, 768), torch.ones(n, S2, 768))]}
    assert torch.equal(O.forward_retrieval(rsd, fam, rb), O.forward_retrieval(rsd, fam, rb, drop=ones2))
