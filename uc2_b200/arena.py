"""Flat parameter / gradient / bf16-shadow arenas in HBM.

The reference keeps ~210 separate fp16 parameter tensors plus apex-amp fp32 masters, copies every
gradient into one flat buffer for the Horovod allreduce (utils/distributed.py:15-42) and copies it
back.  Here the parameters ARE views into one fp32 master arena, the gradients are views into one
fp32 gradient arena (allreduce buckets are plain slices, no copy in/out) and the tensor-core operands
are a parallel bf16 shadow arena refreshed by the fused AdamW kernel.  Module parameter names and
shapes are untouched, so state_dict()/load_state_dict()/named_parameters() behave like the reference.
"""

import torch

from . import _lib

ALIGN = 128   # elements: 512 B in fp32, 256 B in bf16 (TMA needs 16 B, vector kernels 16 B)
ARENAS = []   # most recent arenas; the optimiser finds the one that owns its parameters


def _order(names):
    """Arena order: per BertLayer put query/key/value weights (then biases) back to back so the
    concatenated [2304,768] QKV operand is one contiguous slice."""
    def key(item):
        i, n = item
        if ".attention.self." in n:
            pre = n.split(".attention.self.")[0]
            kind = 0 if n.endswith("weight") else 1
            which = ("query", "key", "value").index(n.split(".attention.self.")[1].split(".")[0])
            return (first[pre], 0, kind, which)
        return (i, 1, 0, 0)
    first = {}
    for i, n in enumerate(names):
        if ".attention.self." in n:
            first.setdefault(n.split(".attention.self.")[0], i)
    return [n for _, n in sorted(enumerate(names), key=key)]


class ParamArena(object):
    def __init__(self, module, device):
        named = [(n, p) for n, p in module.named_parameters()]      # shared params appear once
        self.module = module
        self.device = device
        names = _order([n for n, _ in named])
        params = dict(named)
        self.names = names
        self.index = {n: i for i, n in enumerate(names)}
        self.offset, self.numel, self.shape = {}, {}, {}
        off = 0
        for n in names:
            p = params[n]
            if ".attention.self." in n and not n.endswith("query.weight") and not n.endswith("query.bias"):
                pass                                        # stays glued to the previous tensor of the triple
            else:
                off = (off + ALIGN - 1) // ALIGN * ALIGN
            self.offset[n], self.numel[n], self.shape[n] = off, p.numel(), tuple(p.shape)
            off += p.numel()
        self.total = (off + ALIGN - 1) // ALIGN * ALIGN
        self.master = torch.zeros(self.total, dtype=torch.float32, device=device)
        self.grad = torch.zeros(self.total, dtype=torch.float32, device=device)
        self.shadow = torch.empty(self.total, dtype=torch.bfloat16, device=device)
        with torch.no_grad():
            for n in names:
                p = params[n]
                view = self.master[self.offset[n]:self.offset[n] + p.numel()].view(p.shape)
                view.copy_(p.data.to(device=device, dtype=torch.float32))
                p.data = view
                p.grad = self.grad[self.offset[n]:self.offset[n] + p.numel()].view(p.shape)
        self.params = params
        self.active = {n: False for n in names}             # has this tensor ever received a gradient
        self._versions = None
        self.sync_shadow(force=True)
        ARENAS.append(self)
        del ARENAS[:-8]

    # ------------------------------------------------------------------ views
    def m(self, name):
        return self.params[name].data

    def s(self, name):
        o = self.offset[name]
        return self.shadow[o:o + self.numel[name]].view(self.shape[name])

    def g(self, name):
        o = self.offset[name]
        return self.grad[o:o + self.numel[name]].view(self.shape[name])

    def mp(self, name):
        return self.master.data_ptr() + 4 * self.offset[name]

    def sp(self, name):
        return self.shadow.data_ptr() + 2 * self.offset[name]

    def gp(self, name):
        return self.grad.data_ptr() + 4 * self.offset[name]

    def touch(self, *names):
        for n in names:
            self.active[n] = True

    # ------------------------------------------------------------------ consistency
    def intact(self):
        """False if someone re-allocated parameter storage (e.g. module.to()/half()); then the
        owner rebuilds the arena."""
        n0 = self.names[0]
        return self.params[n0].data_ptr() == self.mp(n0) and self.params[n0].dtype == torch.float32

    def sync_shadow(self, force=False):
        """Re-cast the bf16 shadows if any master tensor was modified through torch (load_state_dict,
        init, a torch optimizer).  The fused AdamW kernel refreshes shadows itself and calls mark_synced()."""
        vers = [self.params[n]._version for n in self.names]
        if force or vers != self._versions:
            _lib.call("uc2_cast_f32_bf16", self.master.data_ptr(), self.shadow.data_ptr(), self.total, _lib.stream())
            self._versions = vers

    def mark_synced(self):
        self._versions = [self.params[n]._version for n in self.names]

    def zero_grad(self):
        self.grad.zero_()
