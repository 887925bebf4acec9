"""Validation loops of the pre-training tasks (pretrain.py:658-1050) around the B200 modules.

Same arithmetic and the same `val_log` keys as the reference.  What changes: the reference pulls three to six
`.item()` scalars to the host per batch (each a full device sync); here every counter is a device scalar (summed in
fp64) and the host reads them once per task.  Cross-rank sums use `distributed.all_gather_list` like the reference.
"""
import time

import torch
import torch.nn.functional as F

from . import distributed as D

IMG_DIM = 2048          # model/const_variable / pretrain.py:51


class _Acc(object):
    """Device-side running sums, read with one host transfer."""

    def __init__(self):
        self.v = {}

    def add(self, **kw):
        for k, x in kw.items():
            x = x.detach().double() if torch.is_tensor(x) else x
            self.v[k] = x if k not in self.v else self.v[k] + x

    def totals(self, *keys):
        vals = [self.v.get(k, 0.0) for k in keys]
        tens = [x for x in vals if torch.is_tensor(x)]
        host = torch.stack(tens).tolist() if tens else []
        it = iter(host)
        local = [next(it) if torch.is_tensor(x) else float(x) for x in vals]
        return [sum(D.all_gather_list(x)) for x in local]


def _soft_correct(out, labels):
    """compute_accuracy_for_soft_targets pretrain.py:989-995."""
    return (out.max(dim=-1)[1] == labels.max(dim=-1)[1]).sum()


@torch.no_grad()
def validate_token_task(model, val_loader, task="mlm"):
    """validate_mlm / validate_mmxlm / validate_vmlm (pretrain.py:815-841, 721-747, 749-775)."""
    acc, st = _Acc(), time.time()
    for batch in val_loader:
        scores = model(batch, task=task, compute_loss=False)
        labels = batch["txt_labels"]
        labels = labels[labels != -1]
        acc.add(loss=F.cross_entropy(scores.float(), labels, reduction="sum"),
                correct=(scores.max(dim=-1)[1] == labels).sum(), n=labels.numel())
    loss, correct, n = acc.totals("loss", "correct", "n")
    tot = time.time() - st
    return {"loss": loss / n, "acc": correct / n, "tok_per_s": n / tot}


def validate_mlm(model, val_loader):
    return validate_token_task(model, val_loader, "mlm")


def validate_mmxlm(model, val_loader):
    return validate_token_task(model, val_loader, "mmxlm")


def validate_vmlm(model, val_loader):
    return validate_token_task(model, val_loader, "vmlm")


@torch.no_grad()
def validate_soft_token_task(model, val_loader, task="mmxlm-soft"):
    """validate_mmxlm_soft / validate_vmlm_soft (pretrain.py:688-719, 777-812)."""
    acc, st = _Acc(), time.time()
    for batch in val_loader:
        pred = F.log_softmax(model(batch, task=task, compute_loss=False).float(), dim=-1)
        tgt = batch["label_targets"]
        acc.add(loss=F.kl_div(pred, tgt, reduction="sum"), score=_soft_correct(pred, tgt),
                n=batch["tgt_masks"].sum())
    loss, score, n = acc.totals("loss", "score", "n")
    tot = time.time() - st
    return {"loss": loss / n, "acc": score / n, "feat_per_s": n / tot}


def validate_mmxlm_soft(model, val_loader):
    return validate_soft_token_task(model, val_loader, "mmxlm-soft")


def validate_vmlm_soft(model, val_loader):
    return validate_soft_token_task(model, val_loader, "vmlm-soft")


@torch.no_grad()
def validate_mrfr(model, val_loader):
    """pretrain.py:881-899."""
    acc, st = _Acc(), time.time()
    for batch in val_loader:
        loss = model(batch, task="mrfr", compute_loss=True)
        acc.add(loss=loss.sum() / IMG_DIM, n=batch["img_mask_tgt"].sum())
    loss, n = acc.totals("loss", "n")
    tot = time.time() - st
    return {"loss": loss / n, "feat_per_s": n / tot}


@torch.no_grad()
def validate_mrc(model, val_loader, task):
    """pretrain.py:947-986 (KL form, or cross entropy against the arg-max class with background excluded)."""
    acc, st = _Acc(), time.time()
    for batch in val_loader:
        pred = model(batch, task=task, compute_loss=False).float()
        tgt = batch["label_targets"]
        if "kl" in task:
            pred = F.log_softmax(pred, dim=-1)
            loss = F.kl_div(pred, tgt, reduction="sum")
            score = _soft_correct(pred, tgt)
        else:
            cls = tgt[:, 1:].max(dim=-1)[1] + 1
            loss = F.cross_entropy(pred, cls, ignore_index=0, reduction="sum")
            score = _soft_correct(pred[:, 1:], tgt[:, 1:])
        acc.add(loss=loss, score=score, n=batch["img_mask_tgt"].sum())
    loss, score, n = acc.totals("loss", "score", "n")
    tot = time.time() - st
    return {"loss": loss / n, "acc": score / n, "feat_per_s": n / tot}


@torch.no_grad()
def validate_itm(model, val_loader):
    """pretrain.py:1004-1050."""
    acc, st = _Acc(), time.time()
    has_ot = False
    for batch in val_loader:
        scores, ot_loss = model(batch, task="itm", compute_loss=False)
        if ot_loss is not None:
            has_ot = True
            if isinstance(ot_loss, tuple):
                pos, neg = ot_loss[0].sum(), ot_loss[1].sum()
                acc.add(ot_pos=pos, ot_neg=neg, ot=pos - neg)
            else:
                acc.add(ot=ot_loss.sum())
        targets = batch["targets"]
        acc.add(loss=F.cross_entropy(scores.float(), targets, reduction="sum"),
                score=(scores.max(dim=-1)[1] == targets).sum(), n=len(targets))
    loss, score, n, ot, ot_pos, ot_neg = acc.totals("loss", "score", "n", "ot", "ot_pos", "ot_neg")
    tot = time.time() - st
    log = {"valid/loss": loss / n, "valid/acc": score / n, "valid/ex_per_s": n / tot}
    if has_ot:
        log["valid/ot_loss"] = ot / n
        log["valid/ot_pos"] = ot_pos / n
        log["valid/ot_neg"] = ot_neg / n
    return log


def validate(model, val_dataloaders, log_fn=None):
    """pretrain.py:658-685: dispatch on the task-name prefix; returns {task: {f'{task}_{k}': v}}."""
    was_training = model.training
    model.eval()
    out = {}
    for task, loader in val_dataloaders.items():
        if task.startswith("mlm"):
            log = validate_mlm(model, loader)
        elif task.startswith("mmxlm-soft"):
            log = validate_mmxlm_soft(model, loader)
        elif task.startswith("mmxlm"):
            log = validate_mmxlm(model, loader)
        elif task.startswith("vmlm-soft"):
            log = validate_vmlm_soft(model, loader)
        elif task.startswith("vmlm"):
            log = validate_vmlm(model, loader)
        elif task.startswith("mrfr"):
            log = validate_mrfr(model, loader)
        elif task.startswith("mrc"):
            log = validate_mrc(model, loader, task)
        elif task.startswith("itm"):
            log = validate_itm(model, loader)
        else:
            raise ValueError(f"Undefined task {task}")
        out[task] = {f"{task}_{k}": v for k, v in log.items()}
        if log_fn is not None:
            log_fn({f"valid_{task}/{k}": v for k, v in out[task].items()})
    if was_training:
        model.train()
    return out
