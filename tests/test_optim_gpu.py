"""GPU: fused multi-tensor AdamW + clip + zero_grad + shadow refresh against the oracle restatement of
optim/adamw.py (itself pinned to the reference's AdamW by tests/golden/adamw.npz), and a short training
run against the oracle training loop."""
import os

import numpy as np
import pytest
import torch

import cases

pytestmark = pytest.mark.gpu


class Opts:
    weight_decay, optim, learning_rate, betas = 0.01, "adamw", 3e-3, (0.9, 0.98)


def _model():
    from uc2_b200 import itm
    cfg = cases.config(1)
    sd = cases.weights(cfg, "retrieval")
    m = itm.VLXLMRForImageTextRetrieval(cfg, 2048)
    m.load_state_dict(sd, strict=False)
    m.cuda().train()
    return m, sd


def test_adamw_kernel_bit_level_semantics():
    from oracle import uc2_oracle as O
    from uc2_b200.optim import build_optimizer, clip_grad_norm_
    m, sd = _model()
    arena = m._arena()
    opt = build_optimizer(m, Opts)
    names = [n for n, _ in m.named_parameters()]
    skipped = {"itm_output.weight", "itm_output.bias"}          # never receive a gradient -> must stay untouched
    ref_p = {n: sd[n].clone() for n in names}
    ref_m = {n: torch.zeros_like(sd[n]) for n in names}
    ref_v = {n: torch.zeros_like(sd[n]) for n in names}
    late = "rank_output.weight"                                  # first gradient only at step 2
    first_step = {}
    for step in range(1, 5):
        lr = Opts.learning_rate * O.warmup_linear(step, 2, 10)
        for g in opt.param_groups:
            g["lr"] = lr
        grads = {}
        for i, n in enumerate(names):
            if n in skipped or (n == late and step == 1):
                continue
            gr = torch.from_numpy(cases.synth.det_normal(sd[n].shape, 900 + 13 * step + i, 0.3))
            grads[n] = gr
            arena.g(n).copy_(gr.cuda())
            arena.touch(n)
            first_step.setdefault(n, step)
        gn = clip_grad_norm_(opt, 5.0)
        opt.step()
        opt.zero_grad()
        total = O.clip_grad_norm(list(grads.values()), 5.0)
        np.testing.assert_allclose(float(gn), float(total), rtol=1e-5)
        for n, gr in grads.items():
            O.adamw_step(ref_p[n], gr, ref_m[n], ref_v[n], step - first_step[n] + 1, lr, 0.9, 0.98, 1e-6,
                         0.0 if O.no_decay(n) else Opts.weight_decay)
        # tensors that have had a gradient before keep stepping with g = 0 (zero_grad leaves zeros, not None)
        for n in names:
            if n not in grads and n in first_step:
                O.adamw_step(ref_p[n], torch.zeros_like(ref_p[n]), ref_m[n], ref_v[n], step - first_step[n] + 1, lr,
                             0.9, 0.98, 1e-6, 0.0 if O.no_decay(n) else Opts.weight_decay)
        assert float(arena.grad.abs().max()) == 0.0              # fused zero_grad
    params = dict(m.named_parameters())
    for n in names:
        np.testing.assert_allclose(params[n].detach().cpu().numpy(), ref_p[n].numpy(), rtol=2e-5, atol=2e-7, err_msg=n)
        np.testing.assert_array_equal(arena.s(n).float().cpu().numpy(),
                                      params[n].detach().to(torch.bfloat16).float().cpu().numpy())
    for n in skipped:
        assert torch.equal(params[n].detach().cpu(), sd[n])


def test_training_steps_follow_oracle():
    """3 optimizer steps of the MRFR task (well conditioned): the loss trajectory tracks the oracle loop."""
    from oracle import uc2_oracle as O
    from uc2_b200 import model
    from uc2_b200.batch import to_device
    from uc2_b200.optim import AdamW
    from uc2_b200.train import TrainStep
    from uc2_b200.utils import set_dropout
    cfg = cases.config(2)
    sd = cases.weights(cfg, "pretrain")
    m = model.VLXLMRForPretraining(cfg, 2048, 1601)
    m.load_state_dict(cases.with_aliases(sd, "pretrain"), strict=False)
    m.cuda().train()
    set_dropout(m, 0)
    lr, wd = 2e-4, 0.01
    decay = [p for n, p in m.named_parameters() if not O.no_decay(n)]
    nodecay = [p for n, p in m.named_parameters() if O.no_decay(n)]
    opt = AdamW([{"params": decay, "weight_decay": wd}, {"params": nodecay, "weight_decay": 0.0}], lr=lr,
                betas=(0.9, 0.98))
    step = TrainStep(m, opt, grad_norm=5.0)
    b = cases.batch_mrfr(seed=31)
    bd = to_device(b, "cuda")
    ref = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    rm = {k: torch.zeros_like(v) for k, v in sd.items()}
    rv = {k: torch.zeros_like(v) for k, v in sd.items()}
    fam = O.Family("vlxlmr")
    got_losses, ref_losses = [], []
    for s in range(1, 4):
        got_losses.append(float(step(bd, "mrfr")))
        for p in ref.values():
            p.grad = None
        l = O.forward_pretraining(ref, fam, b, "mrfr").mean()
        l.backward()
        ref_losses.append(float(l))
        with torch.no_grad():
            ks = [k for k, p in ref.items() if p.grad is not None]
            O.clip_grad_norm([ref[k].grad for k in ks], 5.0)
            for k in ks:
                O.adamw_step(ref[k], ref[k].grad, rm[k], rv[k], s, lr, 0.9, 0.98, 1e-6, 0.0 if O.no_decay(k) else wd)
    np.testing.assert_allclose(got_losses, ref_losses, rtol=2e-3)
    assert ref_losses[2] < ref_losses[0]        # and it actually trains


def test_checkpoint_resume_is_exact(tmp_path):
    """Train 2 steps, save with the reference's checkpoint layout (utils/save.py), restore into a FRESH model and
    optimizer, train 2 more: parameters equal an uninterrupted 4-step run (dropout off; the only run-to-run
    difference is the order of fp32 atomics in the wgrad / embedding-gradient accumulation)."""
    from uc2_b200 import model
    from uc2_b200.batch import to_device
    from uc2_b200.optim import AdamW
    from uc2_b200.save import ModelSaver, TrainingRestorer
    from uc2_b200.train import TrainStep
    from uc2_b200.utils import set_dropout
    from oracle import uc2_oracle as O
    cfg = cases.config(2)
    sd = cases.weights(cfg, "pretrain")
    bd = to_device(cases.batch_mrfr(seed=31), "cuda")

    def fresh():
        m = model.VLXLMRForPretraining(cfg, 2048, 1601)
        m.load_state_dict(cases.with_aliases(sd, "pretrain"), strict=False)
        m.cuda().train()
        set_dropout(m, 0)
        decay = [p for n, p in m.named_parameters() if not O.no_decay(n)]
        nodecay = [p for n, p in m.named_parameters() if O.no_decay(n)]
        opt = AdamW([{"params": decay, "weight_decay": 0.01}, {"params": nodecay, "weight_decay": 0.0}], lr=2e-4,
                    betas=(0.9, 0.98))
        return m, opt, TrainStep(m, opt, grad_norm=5.0)

    m1, o1, s1 = fresh()
    for _ in range(4):
        s1(bd, "mrfr")
    m2, o2, s2 = fresh()
    r2 = TrainingRestorer(str(tmp_path), m2, o2, save_steps=2)
    for _ in range(2):
        s2(bd, "mrfr")
        r2.step()
    ModelSaver(str(tmp_path)).save(m2, 2, o2)
    osd = o2.state_dict()
    assert set(osd) == {"state", "param_groups"} and osd["param_groups"][0]["params"][0] == 0
    n_with_grad = sum(1 for p in m2.parameters() if float(p.grad.abs().sum()) >= 0 and True)
    assert all(st["step"] == 2 for st in osd["state"].values()) and 0 < len(osd["state"]) <= n_with_grad
    assert os.path.exists(tmp_path / "restore.pt") and os.path.exists(tmp_path / "model_step_2.pt")
    assert os.path.exists(tmp_path / "train_state_2.pt")
    m3, o3, s3 = fresh()
    r3 = TrainingRestorer(str(tmp_path), m3, o3, save_steps=2)
    assert r3.global_step == 2 and o3.global_step == 2
    for _ in range(2):
        s3(bd, "mrfr")
    p1, p3 = dict(m1.named_parameters()), dict(m3.named_parameters())
    # Adam divides by sqrt(v): on elements whose gradient is at the fp32-atomics noise floor the update direction is
    # noise too, so two runs may differ there by up to lr = 2e-4 per step.  A resume bug (lost moments, wrong step
    # count, stale shadows) would move EVERY element by that much; here almost none may move.
    worst = []
    for n in p1:
        a, b = p3[n].detach().cpu().numpy(), p1[n].detach().cpu().numpy()
        d = np.abs(a - b)
        worst.append((float(d.max()), float((d > 1e-5).mean()), n))
        assert d.max() <= 4 * 2e-4, (n, float(d.max()))
    frac = max(w[1] for w in worst)
    assert frac < 0.02, sorted(worst, reverse=True)[:5]
    # and the restored run really continued from step 2 (a restart from scratch would sit far away)
    p0 = cases.weights(cfg, "pretrain")
    moved = float((p1["roberta.encoder.layer.0.output.dense.weight"].detach().cpu()
                   - p0["roberta.encoder.layer.0.output.dense.weight"]).abs().mean())
    assert moved > 1e-4


def test_pretrain_loop_end_to_end(tmp_path):
    """PretrainLoop (pretrain.py:484-656) on the real model: MetaLoader task schedule, accumulation 2, validation and
    checkpoints every 2 optimizer steps, meters and throughput counters filled without per-step host reads."""
    from types import SimpleNamespace
    from uc2_b200 import model
    from uc2_b200.batch import to_device
    from uc2_b200.loader import MetaLoader
    from uc2_b200.optim import AdamW
    from uc2_b200.pretrain_loop import PretrainLoop
    from uc2_b200.save import ModelSaver
    from uc2_b200.utils import set_dropout
    cfg = cases.config(2)
    m = model.VLXLMRForPretraining(cfg, 2048, 1601)
    m.load_state_dict(cases.with_aliases(cases.weights(cfg, "pretrain"), "pretrain"), strict=False)
    m.cuda().train()
    set_dropout(m, 0)
    opt = AdamW([{"params": list(m.parameters()), "weight_decay": 0.01}], lr=1e-4, betas=(0.9, 0.98))
    opts = SimpleNamespace(gradient_accumulation_steps=2, num_train_steps=3, valid_steps=2, grad_norm=5.0,
                           itm_ot_lambda=0.1, ot_pos_only=False, learning_rate=1e-4, decay="linear", warmup_steps=1)
    train = {"mlm_coco": [to_device(cases.batch_mlm(seed=s), "cuda") for s in (8, 18)],
             "mrfr_coco": [to_device(cases.batch_mrfr(seed=s), "cuda") for s in (9, 19)],
             "itm_coco": [to_device(cases.batch_itm(seed=s), "cuda") for s in (7, 17)]}
    val = {"mlm_coco": [to_device(cases.batch_mlm(seed=28), "cuda")],
           "itm_coco": [to_device(cases.batch_itm(seed=27), "cuda")]}
    logged, lines = [], []
    loop = PretrainLoop(m, opt, opts, val_dataloaders=val, model_saver=ModelSaver(str(tmp_path)),
                        scalar_log=lambda n, v, s: logged.append((n, v, s)), log=lines.append, log_every=1)
    import random
    random.seed(6)                                                      # schedule: itm, mlm, mrfr
    end = loop.run(MetaLoader(train, accum_steps=2))
    assert end == 3 and opt.global_step == 3
    for s, with_opt in ((2, True), (3, False)):
        assert os.path.exists(tmp_path / f"model_step_{s}.pt")
        assert os.path.exists(tmp_path / f"train_state_{s}.pt") == with_opt
    seen = {k for k, mtr in loop.task2loss.items() if mtr.val is not None}
    assert seen and all(np.isfinite(loop.task2loss[k].val) for k in seen)
    tasks_run = {k for k, v in loop.n_examples.items() if v > 0}
    assert tasks_run == {"itm_coco", "mlm_coco", "mrfr_coco"}
    assert sum(loop.n_examples.values()) == 6 * 6                       # 6 micro-steps of 6 samples
    for t in tasks_run:
        assert loop.n_in_units[t] > 0
        if t.startswith("itm"):
            assert {f"{t}_xe", f"{t}_ot", f"{t}_ot_pos", f"{t}_ot_neg"} <= seen
    names = {n for n, _, _ in logged}
    assert "lr" in names and "grad_norm" in names and any(n.startswith("valid_mlm_coco/") for n in names)
    assert any(n.startswith("valid_itm_coco/itm_coco_valid/") for n in names)
    assert any("examples trained" in l for l in lines)
    assert m.training


def test_deferred_embedding_rows_are_bit_exact():
    """AdamW(lazy_rows=True) postpones the update of vocabulary rows without gradient and replays it when the row is
    needed (csrc/optim.cu, uc2_lazy_table).  Two copies of a model, one eager and one deferred, are fed IDENTICAL
    gradients for nine steps over changing batches and tasks (an MLM step in the middle makes the table dense): every
    forward loss is bit-equal along the way (a forward always sees current rows), rows really are behind in between,
    and after flush() parameters and both moments are bit-identical."""
    from uc2_b200 import model
    from uc2_b200.batch import to_device
    from uc2_b200.optim import AdamW, clip_grad_norm_
    from uc2_b200.train import reduce_loss
    from uc2_b200.utils import set_dropout
    from oracle import uc2_oracle as O
    cfg = cases.config(2)
    sd = cases.weights(cfg, "pretrain")

    def make(lazy):
        m = model.VLXLMRForPretraining(cfg, 2048, 1601)
        m.load_state_dict(cases.with_aliases(sd, "pretrain"), strict=False)
        m.cuda().train()
        set_dropout(m, 0)
        decay = [p for n, p in m.named_parameters() if not O.no_decay(n)]
        nodecay = [p for n, p in m.named_parameters() if O.no_decay(n)]
        opt = AdamW([{"params": decay, "weight_decay": 0.01}, {"params": nodecay, "weight_decay": 0.0}], lr=3e-4,
                    betas=(0.9, 0.98), lazy_rows=lazy)
        return m, opt

    ma, oa = make(False)
    mb, ob = make(True)
    plan = [("mrfr", cases.batch_mrfr(seed=41)), ("itm", cases.batch_itm(seed=42)), ("mrfr", cases.batch_mrfr(seed=43)),
            ("mlm", cases.batch_mlm(seed=44)), ("itm", cases.batch_itm(seed=45)), ("mrc-kl", cases.batch_mrc(seed=46)),
            ("itm", cases.batch_itm(seed=42)), ("mrfr", cases.batch_mrfr(seed=47)), ("itm", cases.batch_itm(seed=48))]
    wname = "roberta.embeddings.word_embeddings.weight"
    behind = 0
    for s, (task, b) in enumerate(plan, 1):
        bd = to_device(b, "cuda")
        la = reduce_loss(ma(bd, task=task, compute_loss=True), task)
        lb = reduce_loss(mb(bd, task=task, compute_loss=True), task)
        assert la.item() == lb.item(), (s, task, la.item(), lb.item())       # same weights where it matters
        la.backward()
        lb.backward()
        aa, ab = ma._arena(), mb._arena()
        ab.grad.copy_(aa.grad)             # identical gradients (the backward's fp32 atomics are not run-to-run exact)
        for o in (oa, ob):
            o.lr_now = 3e-4 * (1.0 - 0.05 * s)
            for g in o.param_groups:
                g["lr"] = o.lr_now         # a changing learning rate: the replay has to use each step's own
            clip_grad_norm_(o, 1.0)
            o.step()
            o.zero_grad()
        torch.cuda.synchronize()
        assert ob._lazy is not None
        if not torch.equal(aa.m(wname), ab.m(wname)):
            behind += 1
    assert behind >= 4, "the deferred path never left a row behind: it is not being exercised"
    ob.flush()
    torch.cuda.synchronize()
    aa, ab = ma._arena(), mb._arena()
    as_bits = lambda t: t.view(torch.int32)
    for what, x, y in (("master", aa.master, ab.master), ("exp_avg", oa.exp_avg, ob.exp_avg),
                       ("exp_avg_sq", oa.exp_avg_sq, ob.exp_avg_sq)):
        bad = (as_bits(x) != as_bits(y)).nonzero().reshape(-1)
        if bad.numel():
            o = aa.offset[wname]
            in_table = int(((bad >= o) & (bad < o + aa.numel[wname])).sum())
            i = int(bad[0])
            print(f"{what}: {bad.numel()} of {x.numel()} elements differ ({in_table} inside the word table); first at {i}: "
                  f"{x[i].item()!r} vs {y[i].item()!r}; table row {(i - o) // 768 if o <= i < o + aa.numel[wname] else None}")
    assert torch.equal(as_bits(aa.master), as_bits(ab.master))
    assert torch.equal(as_bits(oa.exp_avg), as_bits(ob.exp_avg))
    assert torch.equal(as_bits(oa.exp_avg_sq), as_bits(ob.exp_avg_sq))
    assert torch.equal(aa.shadow.view(torch.int16), ab.shadow.view(torch.int16))
    # and the checkpoint view flushes by itself
    sa, sb = ma.state_dict(), mb.state_dict()
    assert all(torch.equal(sa[k], sb[k]) for k in sa)
