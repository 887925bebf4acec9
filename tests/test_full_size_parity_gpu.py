"""GPU: parity against the oracle AT BASELINE.json's sizes (uc2-base: 12 layers, hidden 768, XLM-R vocabulary
250 002), one test per config.  The oracle (oracle/uc2_oracle.py, fp32 on the host cores) needs seconds for a forward
at these sizes, so it is the checker here too; tests/test_full_size_gpu.py adds the size-independent properties.

  configs[1]  120 ragged pairs (8..60 tokens, 10..100 regions): every score within 2e-2 of the oracle
  configs[2]  64 x (60 tokens + 100 regions): the loss of each of the four pre-training tasks within 1e-3 relative
  configs[2]  MLM head at V = 250 002 (decoder pitch padded to a multiple of 8, ragged right edge): logits within
              2e-2, every per-token loss within 2e-2, their mean within 1e-3 relative
  configs[3]  a 12-caption x 96-image sub-grid of the retrieval evaluation: top-10 identical to the oracle in both
              directions (see the test for what "identical" means next to bf16 ties)
  configs[4]  VTLM batch (S = 222, TLM position ids): loss within 1e-3 relative
"""
import numpy as np
import pytest
import torch

import cases

pytestmark = pytest.mark.gpu

FULL_VOCAB = 250002
SCORE_TOL = 2e-2
LOSS_RTOL = 1e-3


def _build(kind, layers=12, seed=42, std=0.02):
    from test_model_gpu import build
    from uc2_b200.utils import set_dropout
    cfg = cases.config(layers, vocab=FULL_VOCAB)
    if std != 0.02:
        from uc2_b200 import synth
        from uc2_b200.config import pretraining_shapes, retrieval_shapes
        shapes = pretraining_shapes(cfg) if kind == "pretrain" else retrieval_shapes(cfg)
        orig = cases.weights
        cases.weights = lambda c, k, family="vlxlmr", seed=seed: synth.fill_state_dict(shapes, seed=seed, perturb=True, std=std)
        try:
            m, sd = build(kind, cfg)
        finally:
            cases.weights = orig
    else:
        m, sd = build(kind, cfg)
    set_dropout(m, 0)
    return m, sd


def _dev(b):
    from uc2_b200.batch import to_device
    return to_device(b, "cuda")


def test_cfg2_full_size_scores_vs_oracle():
    from oracle import uc2_oracle as O
    from uc2_b200 import batch as B, synth
    m, sd = _build("retrieval")
    items = synth.make_pairs(120, seed=77)                     # ragged: 8..60 tokens, 10..100 regions
    batch = B.collate_itm_rank(items, 3)
    with torch.no_grad():
        ref = O.forward_retrieval(sd, O.Family("vlxlmr"), batch, compute_loss=False).numpy().reshape(-1)
        got = m(_dev(batch), compute_loss=False).float().cpu().numpy().reshape(-1)
    err = np.abs(got - ref)
    print(f"cfg2 120 pairs: max |score - oracle| {err.max():.4f}, mean {err.mean():.5f}, score std {ref.std():.4f}")
    assert got.shape == (120,) and err.max() <= SCORE_TOL
    # the triplet loss of the same batch (model/itm.py:43-53)
    m.train()
    with torch.no_grad():
        lref = float(O.forward_retrieval(sd, O.Family("vlxlmr"), batch).mean())
    lgot = float(m(_dev(batch), compute_loss=True).mean())
    assert abs(lgot - lref) <= max(LOSS_RTOL * abs(lref), 2e-4), (lgot, lref)


@pytest.mark.parametrize("task", ["itm", "mlm", "mrfr", "mrc-kl"])
def test_cfg3_full_size_loss_vs_oracle(task):
    """One 64 x (60 + 100) batch of the pre-training mix through 12 layers: loss as pretrain.py:524-553 reduces it."""
    import bench
    from oracle import uc2_oracle as O
    from uc2_b200.train import reduce_loss
    m, sd = _build("pretrain")
    m.train()
    b = bench.pretrain_batches(seed=1234)[task]
    with torch.no_grad():
        ref = float(O.pretraining_loss(O.forward_pretraining(sd, O.Family("vlxlmr"), b, task), task))
        got = float(reduce_loss(m(_dev(b), task=task, compute_loss=True), task))
    print(f"cfg3 {task}: loss {got:.6f} vs oracle {ref:.6f} (rel {abs(got - ref) / abs(ref):.2e})")
    assert abs(got - ref) <= LOSS_RTOL * abs(ref), (task, got, ref)


def test_mlm_head_full_vocabulary_vs_oracle():
    """A13 at V = 250 002: 8 x 60 tokens, ~70 masked rows through the tied decoder; 2 encoder layers (the head is what
    is under test, and the hidden states feeding it are compared as well)."""
    from oracle import uc2_oracle as O
    m, sd = _build("pretrain", layers=2)
    m.train()
    b = cases.batch_mlm(n=8, seed=31, vocab=FULL_VOCAB, txt_len=60, num_bb=36)
    n = int((b["txt_labels"] != -1).sum())
    assert n >= 48 and int(b["txt_labels"].max()) > 200000      # labels reach the top of the vocabulary
    with torch.no_grad():
        fam = O.Family("vlxlmr")
        ref_scores = O.forward_pretraining(sd, fam, b, "mlm", compute_loss=False).numpy()
        ref_loss = O.forward_pretraining(sd, fam, b, "mlm").numpy()
        got_scores = m(_dev(b), task="mlm", compute_loss=False).float().cpu().numpy()
        got_loss = m(_dev(b), task="mlm", compute_loss=True).float().cpu().numpy()
    assert got_scores.shape == (n, FULL_VOCAB) and got_loss.shape == (n,)
    d = np.abs(got_scores - ref_scores)
    print(f"MLM head V=250002: {n} rows, max |logit - oracle| {d.max():.4f}, mean {d.mean():.5f}; "
          f"max |loss - oracle| {np.abs(got_loss - ref_loss).max():.4f}")
    assert d.max() <= 2e-2
    assert np.abs(got_scores[:, -8:] - ref_scores[:, -8:]).max() <= 2e-2      # the ragged right edge of the last tile
    np.testing.assert_allclose(got_loss, ref_loss, atol=2e-2)
    assert abs(got_loss.mean() - ref_loss.mean()) <= LOSS_RTOL * abs(ref_loss.mean())


def test_cfg5_vtlm_full_size_loss_vs_oracle():
    """BASELINE.json configs[4]: 12 x (2 x 60 tokens + 2 specials + 100 regions), S = 222, TLM position ids
    (data/mlm.py:420-428), 12 layers."""
    import bench
    from oracle import uc2_oracle as O
    m, sd = _build("pretrain")
    m.train()
    b = bench.vtlm_batches(seed=99, n=12)["tlm"]
    assert b["attn_masks"].shape[1] == 222
    with torch.no_grad():
        ref = O.forward_pretraining(sd, O.Family("vlxlmr"), b, "tlm").numpy()
        got = m(_dev(b), task="tlm", compute_loss=True).float().cpu().numpy()
    print(f"cfg5 tlm: loss {got.mean():.6f} vs oracle {ref.mean():.6f}, max per-token diff {np.abs(got - ref).max():.4f}")
    assert got.shape == ref.shape
    np.testing.assert_allclose(got, ref, atol=2e-2)
    assert abs(got.mean() - ref.mean()) <= LOSS_RTOL * abs(ref.mean())


def _topk_rows(s, k):
    return np.argsort(-s, axis=1, kind="stable")[:, :k]


def test_cfg4_retrieval_subgrid_top10_vs_oracle():
    """itm.py:516-538 + eval/itm.py:6-53 on a 12-caption x 96-image sub-grid, 12 layers: the top-10 images of every
    caption and the top-10 captions of every image against the oracle's.

    Scores first: within 2e-2 (north_star), and in fact within RANK_ERR = 6e-3 on this grid.  Rankings: on random-init
    weights the 96 image scores of a caption spread with std ~0.07, so neighbours at the rank-10 cut are ~1e-3 apart --
    closer than ANY 16-bit run (this one in bf16, the reference's own in apex fp16) can resolve against an fp32 oracle.
    "Identical" is therefore asserted up to exactly that resolution: the GPU's top-10 may differ from the oracle's only
    by members whose oracle score lies within 2 * RANK_ERR of the oracle's cut (a swap the score bound itself allows),
    at least 8 of every 10 must coincide, and wherever the oracle separates the cut by more than the band the sets
    must be equal.  The counts of exactly equal sets are printed."""
    from oracle import uc2_oracle as O
    from uc2_b200 import batch as B, synth
    RANK_ERR = 6e-3
    m, sd = _build("retrieval")
    n_img, n_cap = 96, 12
    imgs = synth.make_pairs(n_img, seed=501, txt_len=4, bb_range=(10, 100))
    caps = synth.make_pairs(n_cap, seed=502, txt_range=(8, 30), num_bb=10)
    got = np.zeros((n_cap, n_img), np.float32)
    ref = np.zeros((n_cap, n_img), np.float32)
    fam = O.Family("vlxlmr")
    with torch.no_grad():
        for c, cap in enumerate(caps):
            pairs = [dict(input_ids=cap["input_ids"], img_feat=im["img_feat"], img_pos_feat=im["img_pos_feat"]) for im in imgs]
            b = B.collate_itm_rank(pairs, 1)
            ref[c] = O.forward_retrieval(sd, fam, b, compute_loss=False).numpy().reshape(-1)
            got[c] = m(_dev(b), compute_loss=False).float().cpu().numpy().reshape(-1)
    err = np.abs(got - ref)
    print(f"cfg4 sub-grid {n_cap} x {n_img}: max |score - oracle| {err.max():.4f}; oracle score std over images "
          f"{ref.std(1).mean():.3f}, over captions {ref.std(0).mean():.3f}")
    assert err.max() <= RANK_ERR <= SCORE_TOL
    equal = strict = total = 0
    for s_ref, s_got in ((ref, got), (ref.T, got.T)):
        k = 10
        tr, tg = _topk_rows(s_ref, k), _topk_rows(s_got, k)
        for r in range(s_ref.shape[0]):
            total += 1
            srt = np.sort(s_ref[r])[::-1]
            a, b_ = set(tr[r].tolist()), set(tg[r].tolist())
            equal += a == b_
            assert len(a & b_) >= 8, (r, sorted(a), sorted(b_))
            for i in a ^ b_:           # members may only be exchanged inside the resolution band around the cut
                assert abs(s_ref[r, i] - 0.5 * (srt[k - 1] + srt[k])) <= 2 * RANK_ERR, (r, i, s_ref[r, i], srt[k - 1], srt[k])
            if srt[k - 1] - srt[k] > 2 * RANK_ERR:
                strict += 1
                assert a == b_, (r, sorted(a), sorted(b_))
    # top-1 (R@1 of eval/itm.py) wherever the oracle's best is clear of the band
    for s_ref, s_got in ((ref, got), (ref.T, got.T)):
        srt = np.sort(s_ref, 1)[:, ::-1]
        clear = srt[:, 0] - srt[:, 1] > 2 * RANK_ERR
        assert (s_ref.argmax(1)[clear] == s_got.argmax(1)[clear]).all()
    print(f"cfg4 sub-grid: top-10 sets exactly equal in {equal} of {total} rankings ({strict} have an oracle gap above the band)")
