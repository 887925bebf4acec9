"""Optimiser side of the path: AdamW (optim/adamw.py), parameter grouping (optim/misc.py) and LR schedules
(optim/sched.py) with the reference's names and semantics, executed as ONE kernel launch over the flat
parameter arena (uc2_adamw_step) instead of ~10 elementwise launches per parameter tensor.
"""
import ctypes as C
from math import ceil

import torch

from . import _lib
from ._lib import call, stream

CHUNK = 8192


def _find_arena(p):
    from .arena import ARENAS
    for a in reversed(ARENAS):
        lo = a.master.data_ptr()
        if lo <= p.data_ptr() < lo + 4 * a.total and a.intact():
            return a
    raise RuntimeError("uc2_b200.optim.AdamW: parameters are not backed by a uc2_b200 parameter arena yet -- run one "
                       "forward pass on the CUDA device (or call model._arena()) before optimizer.step()")


class AdamW(object):
    """Same constructor and param_groups contract as optim/adamw.py:9-38; step() == adamw.py:40-103.

    Gradient clipping (torch.nn.utils.clip_grad_norm_ at pretrain.py:610 / itm.py:305) is requested with
    ``clip_grad_norm_(optimizer, max_norm)`` below and fused into the same launch; zero_grad() is fused too.
    """

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-6, weight_decay=0.0, correct_bias=True,
                 lazy_rows=False):
        """lazy_rows: postpone the update of word-embedding rows that have no gradient in a step until the row is
        needed (include/uc2_b200.h, uc2_lazy_table).  Parameters and moments end up bit-identical to the eager update;
        every forward pass sees current rows.  Whoever reads the raw parameter tensor or the optimizer state calls
        flush() first (state_dict() of the optimizer and of the models do)."""
        if lr < 0.0:
            raise ValueError("Invalid learning rate: {} - should be >= 0.0".format(lr))
        if not 0.0 <= betas[0] < 1.0:
            raise ValueError("Invalid beta parameter: {} - should be in [0.0, 1.0[".format(betas[0]))
        if not 0.0 <= betas[1] < 1.0:
            raise ValueError("Invalid beta parameter: {} - should be in [0.0, 1.0[".format(betas[1]))
        if not 0.0 <= eps:
            raise ValueError("Invalid epsilon value: {} - should be >= 0.0".format(eps))
        params = list(params)
        if params and not isinstance(params[0], dict):
            params = [{"params": params}]
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, correct_bias=correct_bias)
        self.defaults = defaults
        self.param_groups = []
        for g in params:
            g = dict(g)
            g["params"] = list(g["params"])
            for k, v in defaults.items():
                g.setdefault(k, v)
            self.param_groups.append(g)
        if len(self.param_groups) > 8:
            raise ValueError("at most 8 parameter groups are supported")
        self.global_step = 0
        self._arena = None
        self._pending_clip = 0.0
        self._fused_zero = True
        self.state = {}
        self.lazy_rows = bool(lazy_rows)
        self.lazy_ok = True            # the trainer clears it when several backward passes feed one step
        self._lazy = None              # LazyRows once bound

    # ---------------------------------------------------------------------------------------------
    def _bind(self):
        p0 = self.param_groups[0]["params"][0]
        arena = _find_arena(p0)
        if arena is self._arena:
            return
        if self._arena is not None and self.global_step > 0:
            import warnings
            warnings.warn("uc2_b200.optim.AdamW: the model's parameter arena was rebuilt (model.to() / .half() / a new "
                          "device after the first step); exp_avg / exp_avg_sq restart from zero.  Save and reload the "
                          "optimizer state_dict around such a move to keep the moments.")
        self._arena = arena
        dev = arena.master.device
        by_ptr = {arena.params[n].data_ptr(): n for n in arena.names}
        group_of = [-1] * len(arena.names)
        for gi, g in enumerate(self.param_groups):
            for p in g["params"]:
                n = by_ptr.get(p.data_ptr())
                if n is None:
                    raise RuntimeError("optimizer parameter is not part of the model's arena")
                group_of[arena.index[n]] = gi
        chunks = []
        for n in arena.names:
            t, off, num = arena.index[n], arena.offset[n], arena.numel[n]
            for s in range(0, num, CHUNK):
                chunks.append((off + s, min(CHUNK, num - s), t))
        arr = (_lib.OptChunk * len(chunks))()
        for i, (o, c, t) in enumerate(chunks):
            arr[i].offset, arr[i].n, arr[i].tensor = o, c, t
        raw = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8)
        self._chunks = raw.to(dev)
        self._n_chunks = len(chunks)
        self._group_of = torch.tensor(group_of, dtype=torch.int32, device=dev)
        self._group_of_host = group_of
        self._act_host = [-1] * len(arena.names)
        self._act = torch.full((len(arena.names),), -1, dtype=torch.int32, device=dev)
        self.exp_avg = torch.zeros_like(arena.master)
        self.exp_avg_sq = torch.zeros_like(arena.master)
        self._sqnorm = torch.zeros(1, dtype=torch.float64, device=dev)
        self._lazy = None
        arena.lazy = None
        if self.lazy_rows:
            tables = [n for n in arena.names if n.endswith("embeddings.word_embeddings.weight") and
                      len(arena.shape[n]) == 2 and arena.shape[n][1] % 4 == 0 and group_of[arena.index[n]] >= 0]
            if tables:
                self._lazy = LazyRows(self, arena, tables[0])
                arena.lazy = self._lazy

    def _refresh_active(self, step):
        a = self._arena
        changed = False
        for n in a.names:
            t = a.index[n]
            if a.active[n] and self._act_host[t] < 0 and self._group_of_host[t] >= 0:
                self._act_host[t] = step
                changed = True
        if changed:
            self._act.copy_(torch.tensor(self._act_host, dtype=torch.int32), non_blocking=False)

    def _sparse_step(self):
        """Row ids of the word-embedding gradient when this step can take the deferred path, else None."""
        lz = self._lazy
        if lz is None or not self.lazy_ok:
            return None
        a = self._arena
        if getattr(a, "word_emb_dense", False) or self._act_host[lz.tensor] < 0:
            return None
        if getattr(a, "word_rows_n", 0) != 1:        # no / several embedding backward passes: no complete row list
            return None
        return getattr(a, "word_rows", None)

    def grad_norm(self):
        """Global L2 norm of the gradients of every tensor that has one (device tensor, no host sync)."""
        self._bind()
        self._refresh_active(self.global_step + 1)
        a = self._arena
        rows = self._sparse_step()
        act = self._act if rows is None else self._lazy.act_without_table()
        call("uc2_grad_sqnorm", a.grad.data_ptr(), self._chunks.data_ptr(), self._n_chunks, act.data_ptr(),
             self._sqnorm.data_ptr(), stream())
        if rows is not None:
            self._lazy.add_sqnorm(rows, self._sqnorm)
        return self._sqnorm.sqrt().float()

    def step(self, closure=None):
        loss = closure() if closure is not None else None
        self._bind()
        self.global_step += 1
        self._refresh_active(self.global_step)
        a = self._arena
        h = _lib.AdamwHyper()
        g0 = self.param_groups[0]
        for gi, g in enumerate(self.param_groups):
            h.lr[gi] = g["lr"]
            h.weight_decay[gi] = g["weight_decay"]
            if g["betas"] != g0["betas"] or g["eps"] != g0["eps"] or g["correct_bias"] != g0["correct_bias"]:
                raise NotImplementedError("per-group betas/eps/correct_bias are not used by the reference")
        h.beta1, h.beta2, h.eps = g0["betas"][0], g0["betas"][1], g0["eps"]
        h.correct_bias, h.global_step = int(g0["correct_bias"]), self.global_step
        h.max_grad_norm = float(self._pending_clip)
        h.zero_grad = int(self._fused_zero)
        sq = self._sqnorm.data_ptr() if self._pending_clip > 0 else None
        lz = self._lazy
        rows = self._sparse_step()
        act = self._act
        if lz is not None:
            if rows is None:
                lz.catch_up_all()                    # eager step on the table: every row has to be current first
            else:
                act = lz.act_without_table()
        call("uc2_adamw_step", a.master.data_ptr(), a.grad.data_ptr(), self.exp_avg.data_ptr(),
             self.exp_avg_sq.data_ptr(), a.shadow.data_ptr(), self._chunks.data_ptr(), self._n_chunks,
             act.data_ptr(), self._group_of.data_ptr(), C.byref(h), sq, stream())
        if lz is not None:
            if rows is None:
                lz.mark_all(self.global_step)
            else:
                lz.step_rows(rows, self.global_step, h, sq)
        a.word_rows, a.word_rows_n = None, 0
        a.word_emb_dense = False             # these describe the gradient accumulated for ONE optimizer step
        a.word_dense_sent = False
        self._pending_clip = 0.0
        a.mark_synced()
        return loss

    def flush(self):
        """Apply every postponed row update (no-op without lazy_rows): afterwards the parameter tensors and the
        optimizer state are what eager AdamW would hold."""
        if self._lazy is not None:
            self._lazy.catch_up_all()

    def zero_grad(self, set_to_none=False):
        """Gradients were already cleared inside step() (fused); tensors that never had a gradient are zero."""
        if not self._fused_zero and self._arena is not None:
            self._arena.zero_grad()

    # ---------------------------------------------------------------------------------------------
    # checkpoint format: torch.optim.Optimizer.state_dict() layout, i.e. what the reference's ModelSaver /
    # TrainingRestorer write and read (utils/save.py:58-80, 164-213): {'state': {param index: {'step', 'exp_avg',
    # 'exp_avg_sq'}}, 'param_groups': [{..., 'params': [indices]}]}.  Parameters that never had a gradient have no
    # state entry (optim/adamw.py:52-53 skips them).
    def _param_order(self):
        return [p for g in self.param_groups for p in g["params"]]

    def state_dict(self):
        self.flush()
        groups, k = [], 0
        for g in self.param_groups:
            d = {key: v for key, v in g.items() if key != "params"}
            d["params"] = list(range(k, k + len(g["params"])))
            k += len(g["params"])
            groups.append(d)
        state = {}
        if self._arena is not None:
            a = self._arena
            by_ptr = {a.params[n].data_ptr(): n for n in a.names}
            for i, p in enumerate(self._param_order()):
                n = by_ptr[p.data_ptr()]
                act = self._act_host[a.index[n]]
                if act < 0:
                    continue
                o, num = a.offset[n], a.numel[n]
                state[i] = {"step": self.global_step - act + 1,
                            "exp_avg": self.exp_avg[o:o + num].view(a.shape[n]).clone(),
                            "exp_avg_sq": self.exp_avg_sq[o:o + num].view(a.shape[n]).clone()}
        return {"state": state, "param_groups": groups}

    def load_state_dict(self, sd):
        """Accepts the layout above (from this class or from the reference's optim/adamw.py via torch.optim).
        One incompatibility: the reference's pre-training model registers a `vis_cls.*` head that never receives a
        gradient (model/model.py:468); this build omits it, so a REFERENCE optimizer checkpoint of that model has larger
        parameter groups and a shifted index space and is rejected below with the group-size error -- remap it by
        parameter name first (drop the `vis_cls` slots) if such a checkpoint has to be resumed."""
        self._bind()
        a = self._arena
        if len(sd["param_groups"]) != len(self.param_groups):
            raise ValueError("loaded state dict has a different number of parameter groups")
        for g, saved in zip(self.param_groups, sd["param_groups"]):
            if len(saved["params"]) != len(g["params"]):
                raise ValueError("loaded state dict contains a parameter group that doesn't match the size of "
                                 "optimizer's group")
            for key, v in saved.items():
                if key != "params":
                    g[key] = tuple(v) if key == "betas" else v
        by_ptr = {a.params[n].data_ptr(): n for n in a.names}
        order = self._param_order()
        steps = [int(st["step"]) for st in sd["state"].values()]
        self.global_step = max(steps) if steps else 0
        self.exp_avg.zero_()
        self.exp_avg_sq.zero_()
        self._act_host = [-1] * len(a.names)
        for i, st in sd["state"].items():
            n = by_ptr[order[int(i)].data_ptr()]
            o, num = a.offset[n], a.numel[n]
            self.exp_avg[o:o + num].copy_(st["exp_avg"].reshape(-1).to(self.exp_avg.device, torch.float32))
            self.exp_avg_sq[o:o + num].copy_(st["exp_avg_sq"].reshape(-1).to(self.exp_avg.device, torch.float32))
            self._act_host[a.index[n]] = self.global_step - int(st["step"]) + 1
            a.active[n] = True
        self._act.copy_(torch.tensor(self._act_host, dtype=torch.int32))
        if self._lazy is not None:
            self._lazy.reset(self.global_step)


class LazyRows(object):
    """Host side of the deferred word-embedding update (include/uc2_b200.h uc2_lazy_table; csrc/optim.cu)."""
    HIST = 4096

    def __init__(self, opt, arena, name):
        self.opt, self.arena, self.name = opt, arena, name
        dev = arena.master.device
        self.tensor = arena.index[name]
        self.n_rows, self.width = arena.shape[name]
        self.row_step = torch.full((self.n_rows,), opt.global_step, dtype=torch.int32, device=dev)
        self.row_seen = torch.full((self.n_rows,), -1, dtype=torch.int32, device=dev)
        self.hist = torch.zeros((self.HIST, 2), dtype=torch.float32, device=dev)
        self.pending = False                 # some row may be behind
        self.last_full = opt.global_step     # last step at which every row was current
        self._mark = 0
        self._act_nt = None
        self._act_nt_src = None
        g = opt.param_groups[opt._group_of_host[self.tensor]]
        self.group = opt._group_of_host[self.tensor]
        t = _lib.LazyTable()
        t.param, t.grad = arena.master.data_ptr(), arena.grad.data_ptr()
        t.exp_avg, t.exp_avg_sq = opt.exp_avg.data_ptr(), opt.exp_avg_sq.data_ptr()
        t.table_off, t.n_rows, t.width = arena.offset[name], self.n_rows, self.width
        t.row_step, t.row_seen = self.row_step.data_ptr(), self.row_seen.data_ptr()
        t.hist, t.hist_len = self.hist.data_ptr(), self.HIST
        t.beta1, t.beta2, t.eps = g["betas"][0], g["betas"][1], g["eps"]
        t.decay_on = int(g["weight_decay"] > 0)
        t.shadow_bf16 = arena.shadow.data_ptr()
        self.table = t

    def reset(self, step):
        self.row_step.fill_(step)
        self.pending, self.last_full = False, step

    def act_without_table(self):
        """The optimizer's first-active-step vector with the table switched off (the eager kernel skips it)."""
        if self._act_nt is None or self._act_nt_src != self.opt._act_host:
            host = list(self.opt._act_host)
            host[self.tensor] = -1
            self._act_nt = torch.tensor(host, dtype=torch.int32).to(self.opt._act.device)
            self._act_nt_src = list(self.opt._act_host)
        return self._act_nt

    def _ids(self, rows):
        return rows.reshape(-1).to(torch.long).contiguous()

    def add_sqnorm(self, rows, out):
        ids = self._ids(rows)
        self._mark += 1
        call("uc2_grad_sqnorm_rows", C.byref(self.table), ids.data_ptr(), ids.numel(), self._mark, out.data_ptr(), stream())

    def step_rows(self, rows, step, h, sqnorm_ptr):
        g = self.opt.param_groups[self.group]
        first = self.opt._act_host[self.tensor]
        ids = self._ids(rows)
        call("uc2_adamw_lazy_rows", C.byref(self.table), ids.data_ptr(), ids.numel(), step, first, g["lr"],
             g["weight_decay"], int(g["correct_bias"]), float(h.max_grad_norm), sqnorm_ptr, stream())
        self.pending = True
        if step - self.last_full >= self.HIST - 2:      # the ring is about to wrap over a step some row still needs
            self.catch_up_all()

    def catch_up(self, rows):
        """Before a forward pass reads these rows of the fp32 table."""
        if self.pending and self.opt.global_step > self.last_full:
            ids = self._ids(rows)
            call("uc2_adamw_lazy_catchup", C.byref(self.table), ids.data_ptr(), ids.numel(), self.opt.global_step, stream())

    def catch_up_all(self):
        """Before anything reads the whole table (the tied MLM decoder, checkpoints, an eager step on the table)."""
        if self.pending and self.opt.global_step > self.last_full:
            call("uc2_adamw_lazy_catchup", C.byref(self.table), None, 0, self.opt.global_step, stream())
        self.pending, self.last_full = False, self.opt.global_step

    def mark_all(self, step):
        self.row_step.fill_(step)
        self.pending, self.last_full = False, step


def clip_grad_norm_(optimizer, max_norm):
    """pretrain.py:610 ``clip_grad_norm_(amp.master_params(optimizer), opts.grad_norm)``: returns the total
    norm (device tensor) and arms the clip coefficient for the next optimizer.step()."""
    total = optimizer.grad_norm()
    optimizer._pending_clip = float(max_norm)
    return total


# ---- optim/misc.py:9-32 -------------------------------------------------------------------------------
def build_optimizer(model, opts):
    param_optimizer = list(model.named_parameters())
    no_decay = ["bias", "LayerNorm.bias", "LayerNorm.weight"]
    optimizer_grouped_parameters = [
        {"params": [p for n, p in param_optimizer if not any(nd in n for nd in no_decay)],
         "weight_decay": opts.weight_decay},
        {"params": [p for n, p in param_optimizer if any(nd in n for nd in no_decay)], "weight_decay": 0.0}]
    if opts.optim != "adamw":
        raise ValueError("invalid optimizer (the B200 path implements adamw, the optimiser both shipped configs use)")
    # deferred row updates for the word-embedding table (bit-identical to the eager update) unless switched off
    return AdamW(optimizer_grouped_parameters, lr=opts.learning_rate, betas=opts.betas,
                 lazy_rows=bool(getattr(opts, "lazy_embedding_rows", True)))


# ---- optim/sched.py -----------------------------------------------------------------------------------
def noam_schedule(step, warmup_step=4000):
    if step <= warmup_step:
        return step / warmup_step
    return (warmup_step ** 0.5) * (step ** -0.5)


def warmup_linear(step, warmup_step, tot_step):
    if step < warmup_step:
        return step / warmup_step
    return max(0, (tot_step - step) / (tot_step - warmup_step))


def vqa_schedule(step, warmup_interval, decay_interval, decay_start, decay_rate):
    """optim/sched.py:19-31 (the MCAN schedule: three warm-up plateaus, then stepwise decay)."""
    if step < 3 * warmup_interval:
        return (step // warmup_interval + 1) / 4
    if step >= decay_start:
        return decay_rate ** ceil((step - decay_start) / decay_interval)
    return 1


def _sched(base_lr, global_step, opts):
    if opts.decay == "linear":
        lr = base_lr * warmup_linear(global_step, opts.warmup_steps, opts.num_train_steps)
    elif opts.decay == "invsqrt":
        lr = base_lr * noam_schedule(global_step, opts.warmup_steps)
    elif opts.decay == "constant":
        lr = base_lr
    elif opts.decay == "vqa":
        lr = base_lr * vqa_schedule(global_step, opts.warm_int, opts.decay_int, opts.decay_st, opts.decay_rate)
    else:
        raise ValueError("unsupported decay")
    return lr if lr > 0 else 1e-8         # guard against a miscounted num_train_steps (sched.py:48-50)


def get_lr_sched(global_step, opts):
    """optim/sched.py:34-51."""
    return _sched(opts.learning_rate, global_step, opts)


def get_xlmr_lr_sched(global_step, opts):
    """optim/sched.py:53-70: the same schedules on `opts.xlmr_lr` (itm.py --separate_lr: encoder groups)."""
    return _sched(opts.xlmr_lr, global_step, opts)
