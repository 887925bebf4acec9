"""GPU: fused multi-tensor AdamW + clip + zero_grad + shadow refresh against the oracle restatement of
optim/adamw.py (itself pinned to the reference's AdamW by tests/golden/adamw.npz), and a short training
run against the oracle training loop."""
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
