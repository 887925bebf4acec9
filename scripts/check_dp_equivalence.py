#!/usr/bin/env python
"""Run under torchrun with 2+ GPUs: one ITM fine-tuning backward pass with the row-sparse word-embedding gradient
exchange and one with the dense bucketed all-reduce must leave the SAME averaged gradient arena on every rank, and
that arena must equal the mean of the per-rank single-GPU gradients.  Same for an MLM pre-training backward pass,
whose vocabulary-table gradient has a dense term (the tied decoder): exchanged late in one piece, or early with the
lookup's rows sent separately (GradSync.side_rows), in fp32 and with the 16-bit wire type.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 scripts/check_dp_equivalence.py
"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import cases  # noqa: E402
from uc2_b200 import distributed as D, itm  # noqa: E402
from uc2_b200.batch import to_device  # noqa: E402
from uc2_b200.utils import set_dropout  # noqa: E402


def grads(model, batch, mode, task=None, comm_dtype=None):
    arena = model._arena()
    arena.grad.zero_()
    arena.word_emb_dense = arena.word_dense_sent = False
    sync = None
    if mode != "local":
        sync = D.GradSync(arena.grad, 8 << 20, comm_dtype=comm_dtype)
        sync.layers_per_segment = 1
        sync.allow_sparse = mode == "sparse"
        sync.early_dense = mode == "early"
        arena.grad_sync = sync
    else:
        arena.grad_sync = None
    out = model(batch, task=task, compute_loss=True) if task else model(batch, compute_loss=True)
    out.mean().backward()
    if sync is not None:
        sync.finish()
    torch.cuda.synchronize()
    return arena.grad.clone()


def main():
    D.init("nccl")
    r, w = D.rank(), D.size()
    torch.cuda.set_device(D.local_rank())
    cfg = cases.config(2)
    m = itm.VLXLMRForImageTextRetrieval(cfg, 2048, margin=0.2)
    m.load_state_dict(cases.weights(cfg, "retrieval"), strict=False)
    m.cuda().train()
    set_dropout(m, 0)
    b = to_device(cases.batch_rank(n=12, sample_size=3, seed=100 + r), "cuda")
    g_local = grads(m, b, "local")
    g_dense = grads(m, b, "dense")
    g_sparse = grads(m, b, "sparse")
    ref = g_local.clone()
    dist.all_reduce(ref)
    ref /= w
    e_dense = float((g_dense - ref).abs().max() / ref.abs().max())
    e_sparse = float((g_sparse - ref).abs().max() / ref.abs().max())
    nz = int((g_sparse[:cfg.vocab_size * 768].view(-1, 768).abs().sum(1) > 0).sum())
    print(f"rank {r}: dense vs mean-of-local {e_dense:.2e}, sparse vs mean-of-local {e_sparse:.2e}, "
          f"non-zero vocabulary rows {nz}", flush=True)
    # bf16 forward noise differs run to run only through atomics ordering: ~1e-3 relative is the fp32-atomic floor
    assert e_dense < 5e-3 and e_sparse < 5e-3
    # MLM step of the pre-training model: dense table gradient
    from uc2_b200 import model as umodel
    pm = umodel.VLXLMRForPretraining(cfg, 2048, 1601)
    pm.load_state_dict(cases.with_aliases(cases.weights(cfg, "pretrain"), "pretrain"), strict=False)
    pm.cuda().train()
    set_dropout(pm, 0)
    pb = to_device(cases.batch_mlm(n=8, seed=200 + r), "cuda")
    p_local = grads(pm, pb, "local", "mlm")
    ref = p_local.clone()
    dist.all_reduce(ref)
    ref /= w
    for mode, dt in (("dense", None), ("early", None), ("early", torch.bfloat16)):
        g = grads(pm, pb, mode, "mlm", dt)
        e = float((g - ref).abs().max() / ref.abs().max())
        print(f"rank {r}: mlm {mode}{' bf16 wire' if dt else ''} vs mean-of-local {e:.2e}", flush=True)
        assert e < (2e-2 if dt else 5e-3)
        assert float(pm._arena().word_side.abs().max()) == 0.0 if mode == "early" else True
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
