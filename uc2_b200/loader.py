"""Batch scheduling around the hot path: token-budget length bucketing and the multi-task loader.

Reference behaviour followed (file:line under /root/reference):
  TokenBucketSampler   data/sampler.py:11-59    shuffle ids, cut into buckets, sort each bucket by length (longest
                                                first), fill batches in groups of `size_multiple` samples while
                                                max_len * batch_size stays within the token budget, shuffle batches
  MetaLoader           data/loader.py:13-55     every `accum_steps` micro-steps pick a task from the ratio-weighted
                                                pool; restart a task's iterator when it runs dry; runs forever
  PrefetchLoader       data/loader.py:75-135    -> uc2_b200.batch.Prefetcher

Integer contract: with the same `random` state the sampler yields exactly the reference's batches (the golden
fixture tests/golden/loader.npz was produced by the reference class).  The one deliberate change is in distributed
mode: the reference broadcasts rank 0's task choice through a pickled all-gather every step (a host sync per step,
utils/distributed.py:207-230); here every rank draws the task from its own `random.Random(task_seed)`, so all
ranks agree without talking.
"""
import random


class TokenBucketSampler(object):
    """Yields lists of dataset indices (one list per batch).  `lens[i]` is the padded-relevant length of sample i
    (tokens + regions), `batch_size` the token budget per batch INCLUDING padding."""

    def __init__(self, lens, bucket_size, batch_size, droplast=False, size_multiple=8):
        self._lens = lens
        self._bucket_size = int(bucket_size)
        self._max_tok = int(batch_size)
        self._droplast = bool(droplast)
        self._size_mul = int(size_multiple)

    def _batches_of(self, bucket):
        """Greedy fill of one length-sorted bucket, `size_multiple` samples at a time."""
        out, cur, longest = [], [], 0
        step = self._size_mul
        for s in range(0, len(bucket), step):
            group = bucket[s:s + step]
            longest = max(longest, max(self._lens[i] for i in group))
            if longest * (len(cur) + step) > self._max_tok:
                if not cur:
                    raise ValueError("max_tokens too small / max_seq_len too long")
                out.append(cur)
                cur = list(group)
            else:
                cur = cur + list(group)
        return out, cur

    def __iter__(self):
        ids = list(range(len(self._lens)))
        random.shuffle(ids)
        batches = []
        for s in range(0, len(ids), self._bucket_size):
            bucket = sorted(ids[s:s + self._bucket_size], key=self._lens.__getitem__, reverse=True)
            full, rest = self._batches_of(bucket)
            batches.extend(full)
            if rest and not self._droplast:
                batches.append(rest)
        random.shuffle(batches)
        return iter(batches)

    def __len__(self):
        raise ValueError("NOT supported. This has some randomness across epochs")


class MetaLoader(object):
    """Wraps one loader per task; `loaders` maps task name -> loader or (loader, ratio).  Iterating yields
    (task, batch) forever; the task changes only every `accum_steps` micro-steps so that one optimizer step sees one
    task (different tasks touch different head parameters, pretrain.py:517)."""

    def __init__(self, loaders, accum_steps=1, distributed=False, task_seed=0):
        if not isinstance(loaders, dict):
            raise ValueError("loaders has to be a dict: task -> loader or (loader, ratio)")
        self.name2loader, self.name2iter, self.sampling_pools = {}, {}, []
        for name, entry in loaders.items():
            loader, ratio = entry if isinstance(entry, tuple) else (entry, 1)
            self.name2loader[name] = loader
            self.name2iter[name] = iter(loader)
            self.sampling_pools.extend([name] * int(ratio))
        self.accum_steps = int(accum_steps)
        self.distributed = bool(distributed)
        self.step = 0
        # single process: the global `random` stream, like the reference; several ranks: a private, identically
        # seeded stream per rank replaces the per-step broadcast of rank 0's choice
        self._rng = random.Random(task_seed) if self.distributed else random

    def __iter__(self):
        task = self.sampling_pools[0]
        while True:
            if self.step % self.accum_steps == 0:
                task = self._rng.choice(self.sampling_pools)
            self.step += 1
            try:
                batch = next(self.name2iter[task])
            except StopIteration:
                self.name2iter[task] = iter(self.name2loader[task])
                batch = next(self.name2iter[task])
            yield task, batch
