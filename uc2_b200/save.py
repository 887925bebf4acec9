"""Checkpoint helpers with the reference's file layout (utils/save.py): `ModelSaver` writes
`<prefix>_<step>.pt` (the model state_dict on the CPU, tied aliases included) and `train_state_<step>.pt`
({'step', 'optimizer'}); `TrainingRestorer` keeps `restore.pt` / `restore_backup.pt` with
{'global_step', 'model_state_dict', 'optim_state_dict'} and resumes from whichever loads.  No apex amp state:
the bf16 shadows are derived from the fp32 masters and there is no loss scale."""
import os
from os.path import exists, join

import torch


def _to_cpu(state, half=False):
    """Host copy of a (nested) state.  half=True reproduces the reference's restore.pt compression
    (utils/save.py:147-162: every float tensor stored as fp16); the default keeps fp32 so a resume is exact.
    Either form loads: parameters and optimizer moments are cast back to fp32 on load."""
    if isinstance(state, torch.Tensor):
        t = state.detach().cpu()
        return t.half() if half and t.dtype == torch.float32 else t
    if isinstance(state, dict):
        return {k: _to_cpu(v, half) for k, v in state.items()}
    if isinstance(state, (list, tuple)):
        return type(state)(_to_cpu(v, half) for v in state)
    return state


class ModelSaver(object):
    def __init__(self, output_dir, prefix="model_step", suffix="pt"):
        self.output_dir, self.prefix, self.suffix = output_dir, prefix, suffix

    def save(self, model, step, optimizer=None):
        os.makedirs(self.output_dir, exist_ok=True)
        torch.save(_to_cpu(model.state_dict()), join(self.output_dir, f"{self.prefix}_{step}.{self.suffix}"))
        if optimizer is not None:
            torch.save({"step": step, "optimizer": _to_cpu(optimizer.state_dict())},
                       join(self.output_dir, f"train_state_{step}.pt"))


class TrainingRestorer(object):
    def __init__(self, output_dir, model, optimizer, save_steps=1000, fp16_compress=False):
        self.fp16_compress = fp16_compress
        self.save_path = join(output_dir, "restore.pt")
        self.backup_path = join(output_dir, "restore_backup.pt")
        self.output_dir = output_dir
        self.model, self.optimizer, self.save_steps = model, optimizer, save_steps
        self.global_step = 0
        if exists(self.save_path) or exists(self.backup_path):
            self.restore()

    def step(self):
        self.global_step += 1
        if self.global_step % self.save_steps == 0:
            self.save()

    def save(self):
        os.makedirs(self.output_dir, exist_ok=True)
        checkpoint = {"global_step": self.global_step,
                      "model_state_dict": _to_cpu(self.model.state_dict(), self.fp16_compress),
                      "optim_state_dict": _to_cpu(self.optimizer.state_dict(), self.fp16_compress)}
        if exists(self.save_path):           # keep the previous one in case this write is interrupted
            os.replace(self.save_path, self.backup_path)
        torch.save(checkpoint, self.save_path)

    def restore(self):
        try:
            checkpoint = torch.load(self.save_path, map_location="cpu")
        except Exception:
            checkpoint = torch.load(self.backup_path, map_location="cpu")
        self.global_step = checkpoint["global_step"]
        self.model.load_state_dict(checkpoint["model_state_dict"], strict=False)
        self.model._arena()                   # refresh the bf16 shadows from the loaded masters
        self.optimizer.load_state_dict(checkpoint["optim_state_dict"])


def rename_checkpoint(checkpoint, add_prefix="bert."):
    """pretrain.py:72-80 (--rename_checkpoints): prefix every key of a checkpoint dict, in place."""
    for key in list(checkpoint):
        checkpoint[add_prefix + key] = checkpoint.pop(key)
    return checkpoint


def inject_early_adaptation(checkpoint, early_adaptation_checkpoint, prefix="roberta."):
    """pretrain.py:438-441 (--early_adaptation): the visual-to-word projection learnt in the early-adaptation stage
    becomes the image embedding's input projection."""
    checkpoint[prefix + "img_embeddings.img_linear.weight"] = early_adaptation_checkpoint["v2w_linear.weight"]
    checkpoint[prefix + "img_embeddings.img_linear.bias"] = early_adaptation_checkpoint["v2w_linear.bias"]
    return checkpoint
