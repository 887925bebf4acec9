"""Data-parallel glue: the reference's Horovod helpers (utils/distributed.py) re-expressed over
torch.distributed (NCCL over NVLink 5 / NVSwitch on the GPU box, gloo in CPU tests).

The reference copies every gradient into one flat fp16 buffer AFTER backward, runs one blocking
``hvd.allreduce_`` on ~560 MB and copies back (utils/distributed.py:15-42, called at pretrain.py:564-566).
Here the gradients already live in one contiguous fp32 arena, so a bucket is just a slice: ``GradSync``
launches one asynchronous all-reduce per bucket as soon as the backward pass has finished writing that slice
(heads first, then encoder layers from the top, embeddings last) on NCCL's own stream, overlapping the
remaining backward kernels.  Semantics: mean over ranks (Horovod's default ``average=True`` with
``rescale_denom=1``; SURVEY 8c notes this is unpinned in the reference).
"""
import os

import torch
import torch.distributed as dist


def is_initialized():
    return dist.is_available() and dist.is_initialized()


def size():
    return dist.get_world_size() if is_initialized() else 1


def rank():
    return dist.get_rank() if is_initialized() else 0


def local_rank():
    return int(os.environ.get("LOCAL_RANK", 0))


_CTL = None      # gloo group for host-side control values when the data plane is NCCL


def init(backend=None):
    """One process per GPU (torchrun sets RANK / LOCAL_RANK / WORLD_SIZE / MASTER_*)."""
    global _CTL
    if is_initialized() or int(os.environ.get("WORLD_SIZE", "1")) <= 1:
        return
    backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
    if backend == "nccl":
        torch.cuda.set_device(local_rank())
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank()))
        _CTL = dist.new_group(backend="gloo")
    else:
        dist.init_process_group(backend)


def host_max(value):
    """max over ranks of a host integer, exchanged on the CPU (gloo) so no GPU stream is drained for it."""
    if size() == 1:
        return int(value)
    t = torch.tensor([int(value)], dtype=torch.long)
    if _CTL is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=_CTL)
    elif dist.get_backend() == "gloo":
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    else:                                   # NCCL without a control group: pay one device round trip
        t = t.cuda()
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return int(t.item())


def _avg_inplace(t, async_op=False):
    if t.is_cuda:
        return dist.all_reduce(t, op=dist.ReduceOp.AVG, async_op=async_op)
    w = dist.all_reduce(t, op=dist.ReduceOp.SUM, async_op=async_op)      # gloo has no AVG
    if async_op:
        return _DivAfter(w, t, size())
    t.div_(size())
    return None


class _DivAfter(object):
    def __init__(self, work, t, n):
        self.work, self.t, self.n = work, t, n

    def wait(self):
        self.work.wait()
        self.t.div_(self.n)


class GradSync(object):
    """Bucketed, backward-overlapped gradient averaging over a flat gradient arena.

    comm_dtype (CUDA only): exchange the gradients in this 16-bit type instead of the arena's fp32 -- what the
    reference itself does (apex O2 hands fp16 gradients to all_reduce_and_rescale_tensors, utils/distributed.py:15-42,
    2 bytes per gradient on the wire).  A bucket is cast into a persistent communication buffer, averaged there, and
    cast back into the fp32 arena in finish(); accumulation over ranks happens inside NCCL."""

    def __init__(self, flat_grad, bucket_bytes=64 << 20, comm_dtype=None):
        self.flat = flat_grad
        self.bucket = max(1, bucket_bytes // flat_grad.element_size())
        self.works = []
        self.enabled = True
        self.allow_sparse = False   # row-sparse exchange is only valid when one backward pass feeds the step
        self.done = []          # [lo, hi) ranges already submitted in this step
        self.comm_dtype = comm_dtype if (comm_dtype is not None and flat_grad.is_cuda) else None
        self.comm = None        # persistent 16-bit staging buffer, same indexing as flat
        self.staged = []        # (work, lo, hi) buckets to copy back in finish()
        self.after = []         # closures run by finish() once every bucket is back in the arena
        self.overlap = True     # False: ready() only records the range; everything is exchanged in finish(), after the
                                # backward pass (no NCCL kernel holds SMs while the persistent GEMMs run)
        self.early_dense = True # tied-decoder steps: exchange the dense table gradient as soon as the decoder's wgrad is
                                # written, the embedding lookup's rows separately at the end (side_rows)
        if flat_grad.is_cuda and size() > 1 and "UC2_GEMM_SCHED" not in os.environ:
            # NCCL's kernels will hold SMs while the persistent GEMMs run: let their workers draw tiles from a counter
            # so that a worker that starts late leaves at once (include/uc2_b200.h: uc2_gemm_sched_dynamic)
            from ._lib import lib
            lib().uc2_gemm_sched_dynamic(1)

    def _submit(self, s, e):
        if self.comm_dtype is None:
            self.works.append(_avg_inplace(self.flat[s:e], async_op=True))
            return
        if self.comm is None:
            self.comm = torch.empty(self.flat.numel(), dtype=self.comm_dtype, device=self.flat.device)
        buf = self.comm[s:e]
        buf.copy_(self.flat[s:e])
        self.staged.append((dist.all_reduce(buf, op=dist.ReduceOp.AVG, async_op=True), s, e))

    def ready(self, lo, hi):
        """The backward pass has finished writing flat[lo:hi]."""
        if not self.enabled or size() == 1 or hi <= lo:
            return
        if not self.overlap:
            return                          # finish() submits whatever was not reported
        self.done.append((lo, hi))
        for s in range(lo, hi, self.bucket):
            self._submit(s, min(s + self.bucket, hi))

    def sparse_rows_table(self, lo, n_rows, width, row_ids, pad_row=0):
        """sparse_rows() for a table of exactly n_rows rows starting at flat[lo]."""
        if not self.enabled or size() == 1:
            return
        sub = GradSync.__new__(GradSync)
        sub.flat, sub.bucket, sub.works, sub.enabled, sub.done = self.flat[:lo + n_rows * width], self.bucket, [], True, []
        sub.comm_dtype = self.comm_dtype
        sub.sparse_rows(lo, width, row_ids, pad_row)
        self.last_ids = sub.last_ids
        self.done.append((lo, lo + n_rows * width))

    def sparse_rows(self, lo, width, row_ids, pad_row=0):
        """Exchange a row-sparse slice: flat[lo : lo + rows * width] viewed as [rows, width] whose only non-zero rows
        on this rank are `row_ids` (duplicates allowed).  Used for the word-embedding gradient when no dense term
        (the tied MLM decoder) touched it this step: ITM fine-tuning moves <= B*T rows of the 250 002 x 768 table, so
        every rank all-gathers (ids, rows) -- W x 22 MB at the bench shape -- instead of all-reducing 768 MB of zeros.
        Ranks may hold different numbers of ids (ragged batches): the lists are padded to the largest count (agreed on
        the host through the gloo control group) with `pad_row`, a row that never carries gradient (the embedding's
        padding_idx).  No GPU sync.  Result: mean over ranks, as ready()."""
        if not self.enabled or size() == 1:
            return
        W = size()
        ids = row_ids.reshape(-1).to(torch.long)
        cap = host_max(ids.numel())
        if cap > ids.numel():
            ids = torch.cat([ids, ids.new_full((cap - ids.numel(),), int(pad_row))])
        ids, _ = ids.sort()
        first = torch.ones_like(ids, dtype=torch.bool)
        first[1:] = ids[1:] != ids[:-1]
        n_rows = (self.flat.numel() - lo) // width
        table = self.flat[lo:lo + n_rows * width].view(n_rows, width)
        rows = table.index_select(0, ids) * first.unsqueeze(1).to(table.dtype)     # each local row once
        if getattr(self, "comm_dtype", None) is not None:
            rows = rows.to(self.comm_dtype)
        all_ids = torch.empty(W * ids.numel(), dtype=ids.dtype, device=ids.device)
        all_rows = torch.empty((W * ids.numel(), width), dtype=rows.dtype, device=rows.device)
        dist.all_gather_into_tensor(all_ids, ids)
        dist.all_gather_into_tensor(all_rows, rows)
        table.index_fill_(0, ids, 0)
        table.index_add_(0, all_ids, all_rows.to(table.dtype), alpha=1.0 / W)
        self.last_ids = all_ids                      # every row that carries gradient after the exchange
        self.done.append((lo, lo + n_rows * width))

    def side_rows(self, side, table, row_ids, pad_row=0):
        """Steps whose vocabulary-table gradient has a dense term (the tied MLM decoder's wgrad) AND the embedding
        lookup's row-sparse term: the dense term went out through ready() right after the decoder's backward and is
        being averaged under the encoder's backward; the lookup rows were accumulated into `side` (a zero table of the
        same shape) instead.  Exchange those rows like sparse_rows, clear them in `side`, and add their mean to
        `table` in finish(), after the dense average has landed."""
        W = size()
        ids = row_ids.reshape(-1).to(torch.long)
        cap = host_max(ids.numel())
        if cap > ids.numel():
            ids = torch.cat([ids, ids.new_full((cap - ids.numel(),), int(pad_row))])
        ids, _ = ids.sort()
        first = torch.ones_like(ids, dtype=torch.bool)
        first[1:] = ids[1:] != ids[:-1]
        rows = side.index_select(0, ids) * first.unsqueeze(1).to(side.dtype)
        if self.comm_dtype is not None:
            rows = rows.to(self.comm_dtype)
        all_ids = torch.empty(W * ids.numel(), dtype=ids.dtype, device=ids.device)
        all_rows = torch.empty((W * ids.numel(), rows.size(1)), dtype=rows.dtype, device=rows.device)
        dist.all_gather_into_tensor(all_ids, ids)
        dist.all_gather_into_tensor(all_rows, rows)
        side.index_fill_(0, ids, 0)
        self.after.append(lambda: table.index_add_(0, all_ids, all_rows.to(table.dtype), alpha=1.0 / W))

    def finish(self):
        """Submit whatever was not reported through ready(), then make the compute stream wait."""
        if self.enabled and size() > 1:
            covered = sorted(self.done)
            pos = 0
            for lo, hi in covered + [(self.flat.numel(), self.flat.numel())]:
                if lo > pos:
                    for s in range(pos, lo, self.bucket):
                        self._submit(s, min(s + self.bucket, lo))
                pos = max(pos, hi)
        for w in self.works:
            if w is not None:
                w.wait()
        for w, s, e in self.staged:
            w.wait()
            self.flat[s:e].copy_(self.comm[s:e])
        for fn in self.after:
            fn()
        self.works, self.done, self.staged, self.after = [], [], [], []


# --------------------------------------------------------------------------------------------------
# reference-named helpers (utils/distributed.py)
# --------------------------------------------------------------------------------------------------
def all_reduce_and_rescale_tensors(tensors, rescale_denom):
    """utils/distributed.py:15-42: average every tensor over ranks, then divide by rescale_denom."""
    if size() > 1:
        works = [_avg_inplace(t, async_op=True) for t in tensors]
        for w in works:
            if w is not None:
                w.wait()
    if rescale_denom != 1:
        for t in tensors:
            t.div_(rescale_denom)


def broadcast_tensors(tensors, root_rank, buffer_size=10485760):
    """utils/distributed.py:99-147 (parameter broadcast from rank 0 at start-up)."""
    if size() == 1:
        return
    for t in tensors:
        dist.broadcast(t, root_rank)


def broadcast_arena(arena, root_rank=0):
    """One broadcast of the flat master arena (C2 in SURVEY 2.1), then refresh the bf16 shadows."""
    if size() > 1:
        dist.broadcast(arena.master, root_rank)
        arena.sync_shadow(force=True)


def all_gather_list(data):
    """utils/distributed.py:175-204."""
    if size() == 1:
        return [data]
    out = [None] * size()
    dist.all_gather_object(out, data)
    return out


def any_broadcast(data, root_rank):
    """utils/distributed.py:207-230."""
    if size() == 1:
        return data
    box = [data]
    dist.broadcast_object_list(box, src=root_rank)
    return box[0]


def allgather_rows(t):
    """itm.py:498 ``hvd.allgather(score_matrix)``: concatenate row shards (possibly ragged) over ranks."""
    if size() == 1:
        return t
    n = torch.tensor([t.size(0)], device=t.device)
    ns = [torch.zeros_like(n) for _ in range(size())]
    dist.all_gather(ns, n)
    mx = int(max(int(x) for x in ns))
    pad = t.new_zeros((mx,) + tuple(t.shape[1:]))
    pad[:t.size(0)] = t
    outs = [torch.empty_like(pad) for _ in range(size())]
    dist.all_gather(outs, pad)
    return torch.cat([o[:int(k)] for o, k in zip(outs, ns)], 0)
