"""The pre-training loop of pretrain.py:484-656 around the B200 modules, without horovod / apex / tensorboardX.

Same control flow and the same bookkeeping: one (task name, batch) pair per micro-step from the MetaLoader, the loss
reductions of lines 524-553, an optimizer step every `gradient_accumulation_steps` micro-steps with the LR schedule of
optim/sched.py, per-task RunningMeters (utils/logger.py:71-97), the examples / input-units / loss-units throughput
counters logged every 100 steps, validation + checkpoint every `valid_steps`.

What changes: nothing is read back from the device per micro-step.  The reference calls `.item()` three to six times
a step (loss, xe, ot, ot_pos, ot_neg, the attention-mask sum) and pickles the task name through an all-gather
(pretrain.py:517); here losses and unit counts stay device scalars, queued in order, and are folded into the meters
and counters when something is logged (every `log_every` optimizer steps, default 100, and before validation).  The
meters see the same values in the same order, so their contents at a logging point equal the reference's.
"""
import math
import time
from collections import defaultdict

import torch

from . import distributed as D
from .optim import get_lr_sched
from .train import TrainStep
from .validate import validate as validate_all


class RunningMeter(object):
    """utils/logger.py:71-97: exponentially smoothed scalar, NaN / Inf updates are dropped."""

    def __init__(self, name, val=None, smooth=0.99):
        self._name, self._sm, self._val = name, smooth, val

    def __call__(self, value):
        val = value if self._val is None else value * (1 - self._sm) + self._val * self._sm
        if not math.isnan(val) and not math.isinf(val):
            self._val = val
        else:
            print(f"Inf/Nan in {self._name}")

    def __str__(self):
        return f"{self._name}: {self._val:.4f}"

    @property
    def val(self):
        return self._val

    @property
    def name(self):
        return self._name


class _Deferred(object):
    """Device scalars waiting to be read: (callback, tensor) pairs flushed with ONE host transfer."""

    def __init__(self):
        self.items = []

    def push(self, fn, t):
        self.items.append((fn, t.detach().reshape(()).double()))

    def flush(self):
        from .functional import verify_masked_counts
        verify_masked_counts()          # the same read-back point: the masked-position counts the batches declared
        if not self.items:
            return
        vals = torch.stack([t for _, t in self.items]).tolist()
        for (fn, _), v in zip(self.items, vals):
            fn(v)
        self.items = []


class PretrainLoop(object):
    """`opts` carries the reference's option names: gradient_accumulation_steps, num_train_steps, valid_steps,
    grad_norm, itm_ot_lambda, ot_pos_only, learning_rate, decay, warmup_steps."""

    def __init__(self, model, optimizer, opts, val_dataloaders=None, model_saver=None, restorer=None, log=None,
                 scalar_log=None, log_every=100, step_fn=None):
        self.model, self.optimizer, self.opts = model, optimizer, opts
        self.val_dataloaders = val_dataloaders or {}
        self.model_saver, self.restorer = model_saver, restorer
        self.log = log or (lambda msg: None)
        self.scalar_log = scalar_log or (lambda name, value, step: None)
        self.log_every = log_every
        self.step_fn = step_fn or TrainStep(model, optimizer, grad_norm=opts.grad_norm,
                                            gradient_accumulation_steps=opts.gradient_accumulation_steps,
                                            itm_ot_lambda=opts.itm_ot_lambda, lr_fn=lambda s: get_lr_sched(s, opts))
        self.task2loss = {}
        self.n_examples, self.n_in_units, self.n_loss_units = defaultdict(int), defaultdict(int), defaultdict(int)
        self.pending = _Deferred()
        self.global_step = restorer.global_step if restorer is not None else 0
        self.step_fn.global_step = self.global_step

    # ------------------------------------------------------------------ bookkeeping
    def _meters(self, names):
        o = self.opts
        for name in names:
            self.task2loss.setdefault(name, RunningMeter(f"loss/{name}"))
            if o.itm_ot_lambda > 0 and name.startswith("itm"):
                for suf in ("xe", "ot") + (() if getattr(o, "ot_pos_only", False) else ("ot_pos", "ot_neg")):
                    self.task2loss.setdefault(f"{name}_{suf}", RunningMeter(f"loss/{name}_{suf}"))

    def _record(self, name, task, batch, out, loss):
        """The counters and meters of pretrain.py:518-567, deferred."""
        self.n_examples[name] += batch["input_ids"].size(0)
        units = self.n_in_units
        self.pending.push(lambda v, n=name: units.__setitem__(n, units[n] + int(v)), (batch["attn_masks"] == 1).sum())
        if task.startswith("itm"):
            itm_loss, ot_loss = out
            self.n_loss_units[name] += itm_loss.size(0)
            if ot_loss is not None:
                if not getattr(self.opts, "ot_pos_only", False):
                    pos, neg = ot_loss
                    ot = (pos.sum() - neg.sum()) / (pos.size(0) + neg.size(0))
                    nan_safe = lambda key: (lambda v: None if math.isnan(v) else self.task2loss[key](v))
                    self.pending.push(nan_safe(f"{name}_ot_pos"), pos.mean() if pos.numel() else pos.new_tensor(math.nan))
                    self.pending.push(nan_safe(f"{name}_ot_neg"), neg.mean() if neg.numel() else neg.new_tensor(math.nan))
                else:
                    ot = ot_loss.mean()
                self.pending.push(self.task2loss[f"{name}_xe"], itm_loss.mean())
                self.pending.push(self.task2loss[f"{name}_ot"], ot)
        elif not task.startswith("vmlm-soft"):
            self.n_loss_units[name] += out.size(0)
        self.pending.push(self.task2loss[name], loss)

    def _throughput(self, names, start):
        """pretrain.py:618-641."""
        self.log(f"==============Step {self.global_step}===============")
        for t in names:
            dt = time.time() - start
            tot_ex = sum(D.all_gather_list(self.n_examples[t]))
            tot_in = sum(D.all_gather_list(self.n_in_units[t]))
            tot_l = sum(D.all_gather_list(self.n_loss_units[t]))
            self.log(f"{t}: {tot_ex} examples trained at {int(tot_ex / dt)} ex/s")
            self.scalar_log(f"perf/{t}_ex_per_s", int(tot_ex / dt), self.global_step)
            self.scalar_log(f"perf/{t}_in_per_s", int(tot_in / dt), self.global_step)
            self.scalar_log(f"perf/{t}_loss_per_s", int(tot_l / dt), self.global_step)
        self.log("===============================================")

    def _validate_and_save(self, with_optimizer):
        self.pending.flush()
        self.log(f"Step {self.global_step}: start validation")
        logs = validate_all(self.model, self.val_dataloaders,
                            log_fn=lambda d: [self.scalar_log(k, v, self.global_step) for k, v in d.items()])
        if self.model_saver is not None:
            self.model_saver.save(self.model, self.global_step, self.optimizer if with_optimizer else None)
        return logs

    # ------------------------------------------------------------------ the loop
    def run(self, meta_loader, task_names=None):
        o = self.opts
        names = list(task_names if task_names is not None else getattr(meta_loader, "name2loader", {}).keys())
        self._meters(names)
        start = time.time()
        self.model.train()
        for step, (name, batch) in enumerate(meta_loader):
            if name not in self.task2loss:
                names.append(name)
                self._meters([name])
            task = name.split("_")[0]
            before = self.step_fn.global_step
            loss = self.step_fn(batch, task)
            self._record(name, task, batch, self.step_fn.last_out, loss)
            if self.step_fn.global_step != before:                      # an optimizer step happened
                self.global_step = self.step_fn.global_step
                self.scalar_log("lr", self.optimizer.param_groups[0]["lr"], self.global_step)
                if self.global_step % self.log_every == 0:
                    self.pending.flush()
                    for m in self.task2loss.values():
                        if m.val is not None:
                            self.scalar_log(m.name, m.val, self.global_step)
                    if self.step_fn.last_grad_norm is not None:
                        self.scalar_log("grad_norm", float(self.step_fn.last_grad_norm), self.global_step)
                    self._throughput(names, start)
                if self.global_step % o.valid_steps == 0:
                    self._validate_and_save(with_optimizer=True)
                if self.restorer is not None:
                    self.restorer.step()
            if self.global_step >= o.num_train_steps:
                break
        self.pending.flush()
        if self.global_step % o.valid_steps != 0:
            self._validate_and_save(with_optimizer=False)
        return self.global_step


class FinetuneLoop(object):
    """The retrieval fine-tuning loop of itm.py:253-358: one task, triplet loss, optional separate learning rate for
    the encoder groups (--separate_lr: the first two parameter groups follow `opts.xlmr_lr`), validation (`validate`
    or the full `evaluate`) + checkpoint every `valid_steps`, and a rebuild of the loader after hard negatives were
    re-mined (`steps_per_hard_neg`).

    `opts`: gradient_accumulation_steps, num_train_steps, valid_steps, grad_norm, learning_rate, decay, warmup_steps,
    and optionally separate_lr / xlmr_lr / steps_per_hard_neg.  `build_loader()` returns a fresh iterable of batches
    (the reference re-creates its DataLoader at every pass).  `validate_fn(model) -> dict` and
    `hard_neg_fn(model)` are the caller's closures over their datasets.

    As in PretrainLoop the loss stays on the device between logging points; the reference additionally pickles its
    RunningMeter through an all-gather EVERY optimizer step to average it over ranks (itm.py:297-299) -- here the
    cross-rank mean is taken when the value is logged."""

    def __init__(self, model, optimizer, opts, build_loader, validate_fn=None, hard_neg_fn=None, model_saver=None,
                 restorer=None, log=None, scalar_log=None, log_every=100, step_fn=None):
        from .optim import get_xlmr_lr_sched
        self.model, self.optimizer, self.opts = model, optimizer, opts
        self.build_loader, self.validate_fn, self.hard_neg_fn = build_loader, validate_fn, hard_neg_fn
        self.model_saver, self.restorer = model_saver, restorer
        self.log = log or (lambda msg: None)
        self.scalar_log = scalar_log or (lambda name, value, step: None)
        self.log_every = log_every

        def lr_fn(s):
            lr = get_lr_sched(s, opts)
            if getattr(opts, "separate_lr", False):
                x = get_xlmr_lr_sched(s, opts)
                return [x if i < 2 else lr for i in range(len(optimizer.param_groups))]
            return lr
        self.lr_fn = lr_fn
        self.step_fn = step_fn or TrainStep(model, optimizer, grad_norm=opts.grad_norm,
                                            gradient_accumulation_steps=opts.gradient_accumulation_steps, lr_fn=lr_fn)
        self.running_loss = RunningMeter("loss")
        self.pending = _Deferred()
        self.n_examples = 0
        self.global_step = restorer.global_step if restorer is not None else 0
        self.step_fn.global_step = self.global_step

    def _log_point(self, start):
        self.pending.flush()
        if self.running_loss.val is not None:
            vals = D.all_gather_list(self.running_loss.val)
            self.running_loss = RunningMeter("loss", sum(vals) / len(vals))
            self.scalar_log("loss", self.running_loss.val, self.global_step)
        if self.step_fn.last_grad_norm is not None:
            self.scalar_log("grad_norm", float(self.step_fn.last_grad_norm), self.global_step)
        tot_ex = sum(D.all_gather_list(self.n_examples))
        ex_per_sec = int(tot_ex / (time.time() - start))
        self.log(f"============Step {self.global_step}=============")
        self.log(f"{tot_ex} examples trained at {ex_per_sec} ex/s")
        self.scalar_log("perf/ex_per_s", ex_per_sec, self.global_step)

    def run(self):
        o = self.opts
        hn_every = getattr(o, "steps_per_hard_neg", -1)
        start = time.time()
        self.model.train()
        while True:
            for batch in self.build_loader():
                self.n_examples += batch["input_ids"].size(0)
                before = self.step_fn.global_step
                loss = self.step_fn(batch, None)
                self.pending.push(self.running_loss, loss)
                if self.step_fn.global_step == before:
                    continue                                         # accumulation micro-step
                self.global_step = self.step_fn.global_step
                lrs = self.lr_fn(self.global_step)
                if isinstance(lrs, list):
                    self.scalar_log("xlmr_lr", lrs[0], self.global_step)
                    lrs = lrs[-1]
                self.scalar_log("lr", lrs, self.global_step)
                if self.global_step % self.log_every == 0:
                    self._log_point(start)
                if self.global_step % o.valid_steps == 0 and self.global_step > 0:
                    self.pending.flush()
                    if self.validate_fn is not None:
                        val_log = self.validate_fn(self.model)
                        for k, v in val_log.items():
                            self.scalar_log(k, v, self.global_step)
                    if self.model_saver is not None:
                        self.model_saver.save(self.model, self.global_step)
                if self.restorer is not None:
                    self.restorer.step()
                if hn_every != -1 and self.global_step % hn_every == 0:
                    if self.hard_neg_fn is not None:
                        self.hard_neg_fn(self.model)
                    break                                            # rebuild the loader over the new negatives
                if self.global_step >= o.num_train_steps:
                    break
            if self.global_step >= o.num_train_steps:
                break
        self.pending.flush()
        return self.global_step
