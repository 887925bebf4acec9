"""One optimizer step of the reference training loops around the B200 modules
(pretrain.py:514-650 and itm.py:253-358 without horovod / apex / tensorboard):

    forward -> loss reduction -> backward -> (mean over ranks, overlapped with backward)
            -> lr schedule -> clip_grad_norm_ -> AdamW.step -> zero_grad

Differences from the reference that are deliberate: no per-step pickled all-gathers of the task name
(pretrain.py:517 -- every rank derives the same task from the step number), no .item() host syncs per
micro-step, bf16 + fp32 masters instead of apex O2 loss scaling.
"""

from . import distributed as D
from .optim import clip_grad_norm_


def reduce_loss(out, task, itm_ot_lambda=0.1, ot_pos_only=False):
    """pretrain.py:524-553."""
    if task is None:
        return out.mean()
    if task.startswith("itm"):
        itm_loss, ot_loss = out
        loss = itm_loss.mean()
        if ot_loss is not None:
            if not ot_pos_only:
                ot_pos, ot_neg = ot_loss
                ot = (ot_pos.sum() - ot_neg.sum()) / (ot_pos.size(0) + ot_neg.size(0))
            else:
                ot = ot_loss.mean()
            loss = loss + itm_ot_lambda * ot
        return loss
    if task.startswith("vmlm-soft"):
        return 1000 * out.mean()
    return out.mean()


class TrainStep(object):
    def __init__(self, model, optimizer, grad_norm=-1.0, gradient_accumulation_steps=1, itm_ot_lambda=0.1,
                 lr_fn=None, bucket_bytes=64 << 20, layers_per_segment=3, grad_comm_dtype=None,
                 grad_overlap=True):
        self.model, self.optimizer = model, optimizer
        self.grad_norm = grad_norm
        self.accum = gradient_accumulation_steps
        self.lam = itm_ot_lambda
        self.lr_fn = lr_fn
        self.micro = 0
        self.global_step = 0
        self.bucket_bytes = bucket_bytes
        self.layers_per_segment = layers_per_segment
        self.grad_comm_dtype = grad_comm_dtype      # e.g. torch.bfloat16: 2 bytes per gradient on the wire, as the reference
        self.grad_overlap = grad_overlap            # False: exchange after the backward pass instead of under it
        self.sync = None
        self.trace = None                 # set to [] to collect (start, fwd+bwd done, exchange done, step done) CUDA events
        self.last_out = None              # raw model output of the last micro-step (for loop bookkeeping)
        self.last_grad_norm = None        # device scalar: total gradient norm before clipping, last optimizer step

    def _ensure_sync(self):
        arena = self.model._arena()
        if D.size() > 1 and (self.sync is None or self.sync.flat.data_ptr() != arena.grad.data_ptr()):
            self.sync = D.GradSync(arena.grad, self.bucket_bytes, comm_dtype=self.grad_comm_dtype)
            self.sync.layers_per_segment = self.layers_per_segment if self.grad_overlap else 10 ** 6
            self.sync.overlap = self.grad_overlap
            self.sync.early_dense = self.grad_overlap
            self.sync.allow_sparse = self.accum == 1     # with accumulation earlier micro-steps touched other rows
            arena.grad_sync = self.sync
        return arena

    def __call__(self, batch, task=None):
        """One micro-step; runs the optimizer when the accumulation window closes.  Returns the (device) loss."""
        self._ensure_sync()
        if hasattr(self.optimizer, "lazy_ok"):
            self.optimizer.lazy_ok = self.accum == 1      # several backward passes per step: no complete row list
        last = (self.micro + 1) % self.accum == 0
        if self.sync is not None:
            self.sync.enabled = last               # only the last micro-step's backward triggers communication
        ev = None
        if self.trace is not None:
            import torch
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
            ev[0].record()
        out = self.model(batch, task=task, compute_loss=True) if task is not None else self.model(batch, compute_loss=True)
        loss = reduce_loss(out, task, self.lam, getattr(self.model, "ot_pos_only", False))
        loss.backward()
        if ev:
            ev[1].record()
        self.last_out = out
        self.micro += 1
        if last:
            if self.sync is not None:
                self.sync.finish()
            if ev:
                ev[2].record()
            self.global_step += 1
            if self.lr_fn is not None:
                lr = self.lr_fn(self.global_step)              # one value, or one per parameter group
                groups = self.optimizer.param_groups
                for g, v in zip(groups, lr if isinstance(lr, (list, tuple)) else [lr] * len(groups)):
                    g["lr"] = v
            if self.grad_norm != -1 and self.grad_norm > 0:
                self.last_grad_norm = clip_grad_norm_(self.optimizer, self.grad_norm)
            self.optimizer.step()
            self.optimizer.zero_grad()
            self.model._arena().word_emb_dense = False      # describe the gradient of ONE optimizer step
            self.model._arena().word_dense_sent = False
            if ev:
                ev[3].record()
                self.trace.append((task, ev))
        return loss.detach()
