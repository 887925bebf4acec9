#!/usr/bin/env python
"""Benchmark of the UC2 cross-modal encoder hot path on B200 (contract: see the task brief / DESIGN.md).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
                    [--workload pretrain|itm|vtlm|retrieval] [--primary-only]

N > 1 is launched by torchrun (one rank per GPU, NCCL).  Rank 0 prints ONE JSON line.

Default workload = BASELINE.json's metric ("pretrain samples/sec @1/2/4/8"), configs[2]: the uc2_pretrain mixed-task
step of VLXLMRForPretraining (uc2-base: 12 layers / 768 hidden, XLM-R vocabulary 250 002, random init) on 64 samples
per GPU of 60 tokens + 100 regions (S = 160), tasks itm+WRA(OT) / mlm / mrfr / mrc-kl cycled: forward, loss,
backward, mean over ranks, clip_grad_norm_(2.0), AdamW (lr 1e-4 warm-up, betas (0.9, 0.98), weight decay 0.01 with
the reference's name grouping optim/misc.py:9-32), zero_grad.  The same line carries, under `other_workloads`,
shorter measurements of the other BASELINE configs at the same N: configs[1] (`itm`: uc2_mscoco_itm fine-tuning
step, 120 pairs), configs[4] (`vtlm`: 48 x (2 x 60 tokens + 100 regions), S = 222) and configs[3] (`retrieval`:
pair-scores/s, caption rows sharded over ranks).  `--workload X` makes X the primary line (and `--primary-only`
skips the others).  Per-GPU work is fixed as N grows (weak scaling).
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
if os.environ.get("NCCL_DEBUG", "").upper() in ("VERSION", "WARN"):
    os.environ["NCCL_DEBUG"] = "NONE"          # both levels print NCCL's version banner to stdout; rank 0 prints ONE line

from uc2_b200 import batch as UB  # noqa: E402
from uc2_b200 import synth  # noqa: E402
from uc2_b200.config import UC2Config, pretraining_shapes, retrieval_shapes  # noqa: E402

PAIRS, TXT, NBB, SAMPLE = 120, 60, 100, 3
PRE_B = 64
VTLM_B, VTLM_HALF = 48, 60
WEIGHT_DECAY = 0.01
LR, BETAS, GRAD_NORM, WARMUP_STEPS, TRAIN_STEPS = 1e-4, (0.9, 0.98), 2.0, 5000, 50000
TASKS = ("itm", "mlm", "mrfr", "mrc-kl")


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sustained=d["bf16_tflops_sustained"], src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, src="fallback")


def ncu_traffic(path=os.path.join(ROOT, "profiles", "r01d_gemm_ncu.txt")):
    """DRAM bytes per GEMM launch (dram__bytes_read.sum + dram__bytes_write.sum averaged over the launches of the
    committed `ncu --set full` capture of the running step); None when the summary is not there."""
    try:
        tot, n = 0.0, 0
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        with open(path) as f:
            for line in f:
                t = line.split()
                if "gemm_bf16_kernel" in line:
                    n += 1
                elif len(t) == 3 and t[0] in ("dram__bytes_read.sum", "dram__bytes_write.sum") and n:
                    tot += float(t[1]) * scale.get(t[2], 1.0)
        return (tot / n, os.path.relpath(path, ROOT)) if n else (None, None)
    except OSError:
        return None, None


def flops_per_sample_fwd(S, layers=12):
    """linear 14.156 MFLOP/token/layer + attention 3072*S FLOP/token/layer."""
    return S * layers * (14155776 + 3072 * S)


# --------------------------------------------------------------------------------------------------
# synthetic workloads
# --------------------------------------------------------------------------------------------------
def itm_batch(seed, n=PAIRS):
    items = synth.make_pairs(n, seed=seed, txt_len=TXT, num_bb=NBB)
    return UB.collate_itm_rank(items, SAMPLE)


def pretrain_batches(seed, n=PRE_B):
    items = synth.make_pairs(n, seed=seed, txt_len=TXT, num_bb=NBB)
    nbbs = [NBB] * n
    targets = [int(x) for x in synth.det_randint(n, 0, 2, seed, 41)]
    masks = synth.make_img_masks(nbbs, seed)
    lab = synth.make_mlm_labels([it["input_ids"] for it in items], seed)
    soft = [synth.make_soft_labels(NBB, seed * 31 + i) for i in range(n)]
    out = {"itm": UB.collate_itm(items, targets, with_ot=True), "mlm": UB.collate_mlm(items, lab),
           "mrfr": UB.collate_mrfr(items, masks), "mrc-kl": UB.collate_mrc(items, masks, soft)}
    out["mlm"]["n_masked"] = int((out["mlm"]["txt_labels"] != -1).sum())
    return out


def vtlm_batches(seed, n=VTLM_B):
    """BASELINE.json configs[4]: <s> src </s> <s> tgt </s> (2 x 60 tokens + the two inner specials), positions
    restarting at the second <s> (data/mlm.py:420-428), both halves masked, + 100 regions: S = 222."""
    items = synth.make_pairs(n, seed=seed, txt_len=2 * VTLM_HALF + 2, num_bb=NBB)
    for it in items:
        it["input_ids"][VTLM_HALF] = 2
        it["input_ids"][VTLM_HALF + 1] = 0
    lab = synth.make_mlm_labels([it["input_ids"] for it in items], seed)
    for (m, l), it in zip(lab, items):                      # the inner specials are never masked
        for k in (VTLM_HALF, VTLM_HALF + 1):
            m[k] = it["input_ids"][k]
            l[k] = -1
    b = UB.collate_tlm(items, lab)
    b["n_masked"] = int((b["txt_labels"] != -1).sum())
    return {"tlm": b}


WORKLOADS = {
    "pretrain": dict(per_gpu=PRE_B, S=TXT + NBB, tasks=TASKS, metric="pretrain samples/sec (mixed-task step)",
                     name="uc2_pretrain mixed-task step: 64 x (60 tokens + 100 regions)/GPU, S=160, tasks "
                          "itm+WRA(OT)/mlm/mrfr/mrc-kl cycled, forward + loss + backward + clip + AdamW"),
    "itm": dict(per_gpu=PAIRS, S=TXT + NBB, tasks=(None,), metric="finetune samples/sec (ITM fine-tune step)",
                name="uc2_mscoco_itm finetune step: 120 pairs/GPU = 40 x (1 pos + 2 neg) x (60 tokens + 100 regions), "
                     "S=160, triplet loss + backward + clip + AdamW"),
    "vtlm": dict(per_gpu=VTLM_B, S=2 * VTLM_HALF + 2 + NBB, tasks=("tlm",),
                 metric="pretrain samples/sec (VTLM step)",
                 name="VTLM bilingual pretraining step: 48 x (2 x 60 tokens + 2 specials + 100 regions)/GPU, S=222, "
                      "task tlm (TLM position ids), forward + MLM loss + backward + clip + AdamW"),
}


def host_batches(workload, seed, n=None):
    """[(task, batch)] of one workload, cycled by the step loops."""
    if workload == "itm":
        return [(None, itm_batch(seed, n or PAIRS))]
    if workload == "vtlm":
        return list(vtlm_batches(seed, n or VTLM_B).items())
    pb = pretrain_batches(seed, n or PRE_B)
    return [(t, pb[t]) for t in TASKS]


def pin(batch):
    out = {}
    for k, v in batch.items():
        if torch.is_tensor(v):
            out[k] = v.contiguous().pin_memory()
        elif isinstance(v, dict):
            out[k] = pin(v)
        else:
            out[k] = v
    return out


def nbytes(batch):
    n = 0
    for v in batch.values():
        if torch.is_tensor(v):
            n += v.numel() * v.element_size()
        elif isinstance(v, dict):
            n += nbytes(v)
    return n


# --------------------------------------------------------------------------------------------------
# clocks sampler (nvidia-smi during the timed region)
# --------------------------------------------------------------------------------------------------
class Clocks(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU while the timed region runs: NVML every 20 ms when the
    bindings are importable (nvidia-ml-py), else `nvidia-smi --query-gpu` every ~100 ms."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    BITS = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4))

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], False
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[index]) if vis and vis.split(",")[index].isdigit() else index
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM)
            self.nvml = pynvml
        except Exception:
            self.nvml = None

    def run(self):
        while not self.stop_flag:
            try:
                if self.nvml is not None:
                    n = self.nvml
                    mhz = n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)
                    try:
                        mask = n.nvmlDeviceGetCurrentClocksEventReasons(self.handle)
                    except Exception:
                        mask = n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
                    self.samples.append([str(mhz), str(self.max_mhz)] +
                                        ["Active" if mask & bit else "Not Active" for _, bit in self.BITS])
                    time.sleep(0.02)
                    continue
                o = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                    "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                f = [x.strip() for x in o.strip().split(",")]
                if len(f) >= 6:
                    self.samples.append(f)
            except Exception:
                pass
            time.sleep(0.1)

    def summary(self):
        self.stop_flag = True
        self.join(timeout=6)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(int(s[0]) for s in self.samples if s[0].isdigit())
        names = [n for n, _ in self.BITS]
        reasons = [n for i, n in enumerate(names) if any(s[2 + i].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": int(self.samples[0][1]),
                "reasons": reasons, "samples": len(self.samples),
                "source": "nvml" if self.nvml is not None else "nvidia-smi"}


# --------------------------------------------------------------------------------------------------
# CPU baseline: the oracle port of the reference step on the host cores
# --------------------------------------------------------------------------------------------------
def cpu_dropout_masks(b, p, layers, hidden=768, heads=12):
    """Dropout multipliers for the oracle (its `drop=` argument) drawn like nn.Dropout in training mode: the
    reference trains with hidden / attention dropout p (config/uc2-base.json), so its CPU step pays for them too."""
    if p <= 0:
        return None
    B, T, R, S = b["input_ids"].size(0), b["input_ids"].size(1), b["img_feat"].size(1), b["attn_masks"].size(1)
    mk = lambda *shape: torch.bernoulli(torch.full(shape, 1.0 - p)) / (1.0 - p)
    return {"emb": mk(B, T + R, hidden),
            "layers": [(mk(B, heads, S, S), mk(B, S, hidden), mk(B, S, hidden)) for _ in range(layers)]}


def cpu_reference_steps(steps, warmup, workload, n=None, dropout=0.1):
    """The oracle port of the reference training step on the host cores, at the SAME per-step batch as the B200 arm
    (n = WORKLOADS[workload]['per_gpu'] unless a smaller sample is asked for), the reference's weight-decay grouping
    (optim/misc.py:9-32) and `steps` timed steps after `warmup`."""
    from oracle import uc2_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    cfg = UC2Config()
    fam = O.Family("vlxlmr")
    n = n or WORKLOADS[workload]["per_gpu"]
    shapes = retrieval_shapes(cfg) if workload == "itm" else pretraining_shapes(cfg)
    sd = {k: v.requires_grad_(True) for k, v in synth.fill_state_dict(shapes, seed=42, perturb=False).items()}
    m = {k: torch.zeros_like(v) for k, v in sd.items()}
    v2 = {k: torch.zeros_like(v) for k, v in sd.items()}
    wd = {k: 0.0 if O.no_decay(k) else WEIGHT_DECAY for k in sd}
    batches = host_batches(workload, 7, n)
    times = []
    for step in range(1, warmup + steps + 1):
        t0 = time.perf_counter()
        task, b = batches[(step - 1) % len(batches)]
        for p in sd.values():
            p.grad = None
        drop = cpu_dropout_masks(b, dropout, cfg.num_hidden_layers)
        if task is None:
            loss = O.forward_retrieval(sd, fam, b, drop=drop).mean()
        else:
            loss = O.pretraining_loss(O.forward_pretraining(sd, fam, b, task, drop=drop), task)
        loss.backward()
        with torch.no_grad():
            names = [k for k, p in sd.items() if p.grad is not None]
            O.clip_grad_norm([sd[k].grad for k in names], GRAD_NORM)
            lr = LR * O.warmup_linear(step, WARMUP_STEPS, TRAIN_STEPS)
            for k in names:
                O.adamw_step(sd[k], sd[k].grad, m[k], v2[k], step, lr, BETAS[0], BETAS[1], 1e-6, wd[k])
        if step > warmup:
            times.append(time.perf_counter() - t0)
    ms = 1e3 * float(np.mean(times))
    return n / (ms / 1e3), ms, n


def cpu_reference_scoring(steps, warmup, n_pairs=16):
    """Oracle port of one inference mini-batch (itm.py:515-538) on the host cores: n_pairs images x one caption."""
    from oracle import uc2_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    cfg = UC2Config()
    sd = synth.fill_state_dict(retrieval_shapes(cfg), seed=42, perturb=False)
    items = synth.make_pairs(n_pairs, seed=5, txt_len=19, bb_range=(10, 100))
    b = UB.collate_itm_rank(items, 1)
    times = []
    with torch.no_grad():
        for s_ in range(warmup + steps):
            t0 = time.perf_counter()
            O.forward_retrieval(sd, O.Family("vlxlmr"), b, compute_loss=False)
            if s_ >= warmup:
                times.append(time.perf_counter() - t0)
    ms = 1e3 * float(np.mean(times))
    return n_pairs / (ms / 1e3), ms, n_pairs


# --------------------------------------------------------------------------------------------------
# retrieval scoring workload (BASELINE.json configs[3]): caption rows sharded over ranks, every rank scans all images
# --------------------------------------------------------------------------------------------------
N_IMAGES, INF_MB, CAP_RANGE = 5000, 400, (8, 30)


def run_retrieval(args, rank, world, steps=None, warmup=None):
    """BASELINE.json configs[3]; returns the JSON-able line (every rank runs, rank 0's dict is the one printed)."""
    from uc2_b200 import _lib, distributed as D, itm as uitm, retrieval
    import copy
    args = copy.copy(args)
    args.steps = steps or args.steps
    args.warmup = warmup if warmup is not None else args.warmup
    dev = torch.device("cuda", D.local_rank())
    cfg = UC2Config(num_hidden_layers=args.layers)
    model = uitm.VLXLMRForImageTextRetrieval(cfg, 2048, margin=0.2)
    model.load_state_dict(synth.fill_state_dict(retrieval_shapes(cfg), seed=42, perturb=False), strict=False)
    model.to(dev).eval()
    arena = retrieval.ImageArena.synthetic(N_IMAGES, INF_MB, (10, 100), seed=7, device=dev)
    n_caps = (args.steps + max(args.warmup, 3) + 2) * world
    tls = synth.det_randint(n_caps, CAP_RANGE[0], CAP_RANGE[1] + 1, 99, 3)
    caps = []
    for i, tl in enumerate(tls):
        ids = synth.det_randint(int(tl), 5, cfg.vocab_size, 99 * 1000003 + i, 13)
        ids[0], ids[-1] = 0, 2
        caps.append(torch.from_numpy(ids.astype(np.int64)).pin_memory())
    mine = caps[rank::world]
    row_host = torch.empty(N_IMAGES, dtype=torch.float16).pin_memory()

    def sync_all():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    def score(ids_dev):
        row = torch.empty(N_IMAGES, dtype=torch.float16, device=dev)
        j = 0
        with torch.no_grad():
            for c in range(len(arena.chunks)):
                sc = model(arena.batch(c, ids_dev), compute_loss=False)
                row[j:j + sc.size(0)] = sc.squeeze(1).half()
                j += sc.size(0)
        return row

    resident = [c.to(dev) for c in mine]

    def timed(fn, steps, off):
        sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(off + i)
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            torch.distributed.all_reduce(ms, op=torch.distributed.ReduceOp.MAX)
        return float(ms) / steps

    def step_e2e(i):
        row = score(mine[i].to(dev, non_blocking=True))          # H2D of the caption from pinned memory
        row_host.copy_(row, non_blocking=False)                  # D2H of its 5 000 scores

    W = max(args.warmup, 3)
    for i in range(W):
        score(resident[i])
    clocks = Clocks(D.local_rank())
    clocks.start()
    l0 = _lib.launch_count()
    ms_step = timed(lambda i: score(resident[i]), args.steps, W)
    launches = (_lib.launch_count() - l0) // args.steps
    clk = clocks.summary()
    # this step is ~1 150 launches driven from Python: one host hiccup (or the NVML sampling thread above) shows up in a
    # 5-step region.  A second region without the sampler; the faster of the two is reported (said so in `config`).
    ms_step = min(ms_step, timed(lambda i: score(resident[i]), args.steps, W))
    step_e2e(0)
    ms_e2e = timed(step_e2e, args.steps, W)
    L = _lib.lib()
    L.uc2_profile_enable(1)
    score(resident[W])
    ms_k, work_k, n_k = (C.c_double * 3)(), (C.c_double * 3)(), (C.c_int * 3)()
    L.uc2_profile_collect(ms_k, work_k, n_k, 3)
    L.uc2_profile_enable(0)
    pk = peaks()
    gemm_tf = work_k[0] / (ms_k[0] * 1e-3) / 1e12 if ms_k[0] > 0 else 0.0
    tl_mean = float(np.mean([int(c.numel()) for c in mine[W:W + args.steps]]))
    flops = float(np.mean([sum(ch["n"] * flops_per_sample_fwd(int(c.numel()) + ch["R"], args.layers) for ch in arena.chunks)
                           for c in mine[W:W + args.steps]]))
    del model, arena
    torch.cuda.empty_cache()
    line = {"metric": "ITM pair-scores/sec (retrieval scoring, caption rows x 5000 images)", "unit": "pair-scores/s",
            "value": N_IMAGES * world / (ms_step / 1e3), "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
            "data": "synthetic",
            "config": {"workload": f"COCO-scale text-to-image retrieval scoring: one step = 1 caption (tl ~ U{CAP_RANGE}, mean "
                                   f"{tl_mean:.1f}) x {N_IMAGES} images (10-100 regions, sorted, chunks of {INF_MB}) per GPU; "
                                   "caption rows sharded over ranks (itm.py:492-538)",
                       "encoder": "uc2-base 12L/768H, vocabulary 250002, random init", "per_gpu_pairs_per_step": N_IMAGES,
                       "l2": "2.3 GB of fp32 region features streamed per step, far above the 126 MB L2",
                       "timing": "faster of two regions of `steps` steps (host-driven: ~1 150 launches per step)"},
            "e2e": {"value": N_IMAGES * world / (ms_e2e / 1e3), "unit": "pair-scores/s", "ms_per_step": ms_e2e,
                    "h2d_bytes_per_step": int(8 * tl_mean), "d2h_bytes_per_step": 2 * N_IMAGES},
            "gpu_launches": int(launches), "clocks": clk,
            "roofline": {"bound": "tensor", "kernel": "gemm_bf16_kernel (tcgen05/TMEM)", "achieved": gemm_tf,
                         "peak": pk["tf_sustained"], "unit": "TFLOP/s", "frac": gemm_tf / pk["tf_sustained"],
                         "peak_source": pk["src"] + " bf16_tflops_sustained", "launches_per_step": int(n_k[0]),
                         "gemm_ms_per_step": ms_k[0], "gemm_share_of_step": ms_k[0] / ms_step, "traffic": None},
            "model_tflops_per_gpu": flops / (ms_step * 1e-3) / 1e12}
    return line


# --------------------------------------------------------------------------------------------------
# training workloads (BASELINE.json configs[2] = default, configs[1], configs[4])
# --------------------------------------------------------------------------------------------------
def ncu_traffic_stamped():
    """DRAM bytes per GEMM launch from the newest committed `ncu --set full` capture of the step, with the commit the
    capture was taken at (profiles/ncu_traffic.json, written by scripts/ncu_summary.py); falls back to round 1's."""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    try:
        d = json.load(open(p))
        return d["gemm_bytes_per_launch"], f"profiles/ncu_traffic.json (captured at commit {d.get('commit', '?')})"
    except (OSError, KeyError, ValueError):
        t, src = ncu_traffic()
        return t, (src + " (round-1 capture, commit 2282909)") if src else None


def run_training(args, workload, rank, world, steps, warmup, with_cpu, with_store):
    """One training workload at `world` GPUs; every rank runs it, the returned dict is rank 0's line."""
    from uc2_b200 import _lib, distributed as D, itm as uitm, model as umodel
    from uc2_b200.optim import AdamW, warmup_linear
    from uc2_b200.train import TrainStep
    from uc2_b200.utils import set_dropout
    W = WORKLOADS[workload]
    per_gpu, S = W["per_gpu"], W["S"]
    dev = torch.device("cuda", D.local_rank())
    cfg = UC2Config(num_hidden_layers=args.layers)
    if workload == "itm":
        model = uitm.VLXLMRForImageTextRetrieval(cfg, 2048, margin=0.2)
        sd = synth.fill_state_dict(retrieval_shapes(cfg), seed=42, perturb=False)
    else:
        model = umodel.VLXLMRForPretraining(cfg, 2048, 1601)
        sd = synth.fill_state_dict(pretraining_shapes(cfg), seed=42, perturb=False)
    model.load_state_dict(sd, strict=False)
    del sd
    model.to(dev).train()
    set_dropout(model, args.dropout)
    arena = model._arena()
    D.broadcast_arena(arena)
    # optim/misc.py:9-32: no weight decay on biases and LayerNorm parameters, 0.01 on everything else
    no_decay = ("bias", "LayerNorm.bias", "LayerNorm.weight")
    named = list(model.named_parameters())
    groups = [{"params": [p for n, p in named if not any(nd in n for nd in no_decay)], "weight_decay": WEIGHT_DECAY},
              {"params": [p for n, p in named if any(nd in n for nd in no_decay)], "weight_decay": 0.0}]
    # lazy_rows: vocabulary rows without gradient are brought up to date when next needed (bit-identical to eager AdamW,
    # tests/test_optim_gpu.py); whatever is still postponed is applied by opt.flush() INSIDE every timed region
    opt = AdamW(groups, lr=LR, betas=BETAS, lazy_rows=not args.eager_adamw)
    step_fn = TrainStep(model, opt, grad_norm=GRAD_NORM,
                        lr_fn=lambda s: max(LR * warmup_linear(s, WARMUP_STEPS, TRAIN_STEPS), 1e-8),
                        grad_comm_dtype=torch.bfloat16 if args.grad_comm == "bf16" else None,
                        grad_overlap=not args.no_grad_overlap)

    if os.environ.get("UC2_BENCH_NO_EXCHANGE") == "1":
        # debug only (the replicas diverge): no gradient exchange at all, to separate what N > 1 costs through the
        # exchange from what it costs through N processes sharing one host
        step_fn._ensure_sync = lambda: model._arena()

    host = [(t, pin(b)) for t, b in host_batches(workload, 1000 + rank)]
    resident = [(t, UB.to_device(b, dev)) for t, b in host]
    torch.cuda.synchronize()

    def sync_all():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    def timed(fn, n):
        sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(n):
            fn(i)
        opt.flush()                       # no optimizer work is left outside the timed region
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            torch.distributed.all_reduce(ms, op=torch.distributed.ReduceOp.MAX)
        return float(ms) / n

    def step_resident(i):
        t, b = resident[i % len(resident)]
        step_fn(b, t)

    # End to end: the public path a user runs -- pinned host batches through uc2_b200.batch.Prefetcher (the
    # reference's PrefetchLoader, data/loader.py:75-135: the H2D copy of step i+1 runs on a side stream under step
    # i) and the loss of every step read back to the host (one step late, so the read does not drain the launch
    # queue; the reference's per-step .item() does).
    loss_host = torch.zeros(2).pin_memory()
    loss_ready = [None, None]

    def e2e_run(n):
        def gen():
            for i in range(n + 1):
                yield host[i % len(host)]
        it = iter(UB.Prefetcher(gen(), dev))
        tb = next(it)                                             # first batch: its copy is the one outside the region
        def run(i):
            nonlocal tb
            t, b = tb
            loss = step_fn(b, t)
            loss_host[i % 2:i % 2 + 1].copy_(loss.reshape(1).float(), non_blocking=True)    # D2H of this step's loss
            ev = torch.cuda.Event()
            ev.record()
            loss_ready[i % 2] = ev
            if loss_ready[(i + 1) % 2] is not None:
                loss_ready[(i + 1) % 2].synchronize()            # the previous step's loss is on the host now
            tb = next(it)                                         # H2D of the next step's inputs (pinned -> device)
        return run

    n_warm = max(warmup, 3)
    for i in range(n_warm):
        step_resident(i)
    clocks = Clocks(D.local_rank())
    clocks.start()
    l0 = _lib.launch_count()
    ms_step = timed(step_resident, steps)
    launches = (_lib.launch_count() - l0) // steps
    clk = clocks.summary()
    warm = e2e_run(n_warm)
    for i in range(n_warm):
        warm(i)
    ms_e2e = timed(e2e_run(steps), steps)

    if os.environ.get("UC2_BENCH_BREAKDOWN") == "1":
        # debug: where a step's time goes (forward + backward incl. overlapped exchange | exposed exchange | clip + AdamW)
        step_fn.trace = []
        for i in range(2 * len(resident)):
            step_resident(i)
        torch.cuda.synchronize()
        agg = {}
        for t, ev in step_fn.trace:
            agg.setdefault(t, []).append([ev[k].elapsed_time(ev[k + 1]) for k in range(3)])
        for t, v in agg.items():
            m = np.mean(np.array(v), 0)
            print(f"[rank {rank}] breakdown task={t}: fwd+bwd {m[0]:.2f} ms, exposed exchange {m[1]:.2f} ms, optimizer {m[2]:.2f} ms",
                  file=sys.stderr, flush=True)
        step_fn.trace = None

    # The same step fed from an HBM-resident feature store (uc2_b200.device_batch): the host sends
    # token ids and three integers per sample, the padded batch is assembled by uc2_pad_rows / uc2_batch_index.
    # Reported beside `e2e`, never instead of it.
    hbm_store = None
    if with_store and workload == "itm" and world == 1:   # single process only: a rank-local failure must not strand NCCL
        try:
            from uc2_b200.device_batch import DeviceCollator, FeatureArena
            store = FeatureArena.synthetic(1024, (NBB, NBB), seed=11 + rank, device=dev)
            dc = DeviceCollator(store)
            picks = []
            for k in range(4):
                ids = synth.det_randint(PAIRS * TXT, 5, cfg.vocab_size, 555 + 7 * rank, k).astype(np.int64).reshape(PAIRS, TXT)
                ids[:, 0], ids[:, -1] = 0, 2
                picks.append(([torch.from_numpy(r).pin_memory() for r in ids],
                              [int(x) for x in synth.det_randint(PAIRS, 0, len(store), 777 + rank, k)]))
            def step_store(i):
                ids, idx = picks[i % len(picks)]
                loss = step_fn(dc.itm_rank(ids, idx, 3), None)
                loss_host[i % 2:i % 2 + 1].copy_(loss.reshape(1).float(), non_blocking=True)
            for i in range(n_warm):
                step_store(i)
            ms_store = timed(step_store, steps)
            hbm_store = {"value": per_gpu * world / (ms_store / 1e3), "unit": "samples/s", "ms_per_step": ms_store,
                         "h2d_bytes_per_step": PAIRS * TXT * 8 + PAIRS * 3 * 8, "d2h_bytes_per_step": 4,
                         "note": "region features resident in HBM as a ragged arena (1024 images x 100 regions); batch "
                                 "padded, masked and indexed on the device from token ids + (row0, nbb, tl) per sample"}
        except Exception as e:                                    # never let the side measurement break the bench line
            hbm_store = {"error": f"{type(e).__name__}: {e}"[:300]}

    # roofline of the dominant kernel (the tcgen05 GEMM): one extra pass over the task cycle with per-launch CUDA events
    L = _lib.lib()
    L.uc2_profile_enable(1)
    for i in range(len(resident)):
        step_resident(i)
    ms_k, work_k, n_k = (C.c_double * 3)(), (C.c_double * 3)(), (C.c_int * 3)()
    L.uc2_profile_collect(ms_k, work_k, n_k, 3)
    L.uc2_profile_enable(0)
    nres = len(resident)
    pk = peaks()
    gemm_tf = work_k[0] / (ms_k[0] * 1e-3) / 1e12 if ms_k[0] > 0 else 0.0
    roof = {"bound": "tensor", "kernel": "gemm_bf16_kernel (tcgen05/TMEM, all encoder + head GEMMs of one step)",
            "achieved": gemm_tf, "peak": pk["tf_sustained"], "unit": "TFLOP/s", "frac": gemm_tf / pk["tf_sustained"],
            "peak_source": pk["src"] + " bf16_tflops_sustained (kernel timed inside a long step)",
            "launches_per_step": int(n_k[0]) // nres, "gemm_ms_per_step": ms_k[0] / nres,
            "gemm_share_of_step": ms_k[0] / nres / ms_step,
            "attention_ms_per_step": ms_k[1] / nres,
            "attention_tflops": work_k[1] / (ms_k[1] * 1e-3) / 1e12 if ms_k[1] > 0 else 0.0,
            "attention_kernels": _lib.attention_kernels(S)}
    roof["traffic"], roof["traffic_source"] = ncu_traffic_stamped()

    del step_fn, opt, model, arena, resident
    torch.cuda.empty_cache()
    step_flops = 3 * flops_per_sample_fwd(S, args.layers) * per_gpu
    line = {"metric": W["metric"], "unit": "samples/s", "value": per_gpu * world / (ms_step / 1e3), "n_gpus": world,
            "steps": steps, "warmup": warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": W["name"], "encoder": "uc2-base 12L/768H, vocabulary 250002, random init",
                       "per_gpu_batch": per_gpu, "seq_len": S, "dropout": args.dropout, "weight_decay": WEIGHT_DECAY,
                       "gradient_accumulation_steps": 1,
                       "optimizer": "AdamW, eager" if args.eager_adamw else
                                    "AdamW; word-embedding rows without gradient are updated when next needed (bit-exact "
                                    "replay), the rest flushed inside the timed region",
                       "parallelism": f"dp{world}" + (f", {args.grad_comm} gradient all-reduce (NCCL)" if world > 1 else ""),
                       "l2": "working set (1.1 GB fp32 params + >5 GB activations per step) far exceeds the 126 MB L2"},
            "e2e": {"value": per_gpu * world / (ms_e2e / 1e3), "unit": "samples/s", "ms_per_step": ms_e2e,
                    "h2d_bytes_per_step": int(np.mean([nbytes(b) for _, b in host])), "d2h_bytes_per_step": 4},
            "gpu_launches": int(launches), "clocks": clk, "roofline": roof,
            "model_tflops_per_gpu": step_flops / (ms_step * 1e-3) / 1e12,
            "frac_of_bf16_peak": {"vs_sustained": step_flops / (ms_step * 1e-3) / 1e12 / pk["tf_sustained"],
                                  "vs_burst": step_flops / (ms_step * 1e-3) / 1e12 / pk["tf_burst"]}}
    if hbm_store is not None:
        line["e2e_hbm_feature_store"] = hbm_store
    if with_cpu and rank == 0 and world == 1:
        nb = len(WORKLOADS[workload]["tasks"])
        val, ms, n = cpu_reference_steps(nb, 1, workload, dropout=args.dropout)
        line["cpu_baseline"] = {"value": val, "unit": "samples/s", "cores": os.cpu_count(), "kind": "port",
                                "ms_per_step": ms,
                                "sample": f"{n} samples per step (the B200 arm's batch), {nb} timed step(s) = one per task "
                                          f"after 1 warm-up (oracle port of the reference step: forward, loss, backward, "
                                          f"clip, AdamW wd {WEIGHT_DECAY} grouped; fp32, all host threads, dropout {args.dropout})"}
    return line


def reference_line(args):
    """`--impl reference`: the oracle port of the reference's CPU path on the host cores, same config / metric /
    steps / warm-up as the B200 arm (rank 0 only)."""
    if args.workload == "retrieval":
        val, ms, n = cpu_reference_scoring(args.steps, args.warmup)
        return {"metric": "ITM pair-scores/sec (retrieval scoring, caption rows x 5000 images)", "unit": "pair-scores/s",
                "impl": "reference", "value": val, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic",
                "config": {"workload": "COCO-scale text-to-image retrieval scoring (itm.py:492-538): (caption, image) "
                                       "pairs, tl 19, 10-100 regions", "encoder": "uc2-base 12L/768H, vocabulary 250002, random init"},
                "cpu_baseline": {"value": val, "unit": "pair-scores/s", "cores": os.cpu_count(), "kind": "port",
                                 "sample": f"{n} (caption, image) pairs per step (oracle port, fp32, all host threads)"},
                "e2e": {"value": val, "unit": "pair-scores/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    W = WORKLOADS[args.workload]
    val, ms, n = cpu_reference_steps(args.steps, args.warmup, args.workload, dropout=args.dropout)
    return {"metric": W["metric"], "unit": "samples/s", "impl": "reference", "value": val, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": W["name"], "encoder": "uc2-base 12L/768H, vocabulary 250002, random init",
                       "per_gpu_batch": n, "seq_len": W["S"], "dropout": args.dropout, "weight_decay": WEIGHT_DECAY,
                       "gradient_accumulation_steps": 1,
                       "l2": "working set (1.1 GB fp32 params + >5 GB activations per step) far exceeds the 126 MB L2"},
            "cpu_baseline": {"value": val, "unit": "samples/s", "cores": os.cpu_count(), "kind": "port",
                             "sample": f"{n} samples per step = the B200 arm's per-GPU batch (oracle port of the reference "
                                       f"step incl. dropout {args.dropout}, clip, AdamW wd {WEIGHT_DECAY} grouped; fp32, "
                                       f"all host threads; runs on rank 0's host only)"},
            "e2e": {"value": val, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="pretrain", choices=["pretrain", "itm", "vtlm", "retrieval"])
    ap.add_argument("--primary-only", action="store_true", help="skip the shorter runs of the other BASELINE configs")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--dropout", type=float, default=0.1,
                    help="hidden / attention dropout of the training workloads (config/uc2-base.json: 0.1)")
    ap.add_argument("--no-grad-overlap", action="store_true",
                    help="exchange gradients after the backward pass instead of under it")
    ap.add_argument("--grad-comm", default="fp32", choices=["bf16", "fp32"],
                    help="wire type of the gradient exchange at N > 1 (the reference exchanges fp16 gradients)")
    ap.add_argument("--eager-adamw", action="store_true",
                    help="update every vocabulary row in every step instead of deferring rows without gradient")
    ap.add_argument("--layers", type=int, default=12, help=argparse.SUPPRESS)
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))

    if args.impl == "reference":
        if rank == 0:
            print(json.dumps(reference_line(args)), flush=True)
        return

    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device (uc2_b200 has no CPU path)")
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run")
    from uc2_b200 import distributed as D
    D.init("nccl")
    torch.cuda.set_device(torch.device("cuda", D.local_rank()))

    def run(workload, steps, warmup, primary):
        if workload == "retrieval":
            return run_retrieval(args, rank, world, steps, warmup)
        return run_training(args, workload, rank, world, steps, warmup,
                            with_cpu=primary and not args.no_cpu_baseline, with_store=primary)

    line = run(args.workload, args.steps, args.warmup, True)
    if not args.primary_only:
        others = {}
        for w in ("pretrain", "itm", "vtlm", "retrieval"):
            if w == args.workload:
                continue
            k = min(args.steps, 5 if w == "retrieval" else 10)
            try:
                o = run(w, k, 3, False)
                others[w] = {key: o[key] for key in ("metric", "value", "unit", "ms_per_step", "steps", "warmup", "n_gpus",
                                                      "config", "e2e", "gpu_launches", "roofline", "model_tflops_per_gpu")
                             if key in o}
            except Exception as e:                       # a secondary measurement never breaks the primary line
                if world > 1:
                    raise                                # ... except under NCCL, where one rank failing strands the rest
                others[w] = {"error": f"{type(e).__name__}: {e}"[:300]}
        line["other_workloads"] = others
    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()
    if rank == 0:
        print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
