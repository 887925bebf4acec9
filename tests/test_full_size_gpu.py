"""GPU: BASELINE.json configs[1] at FULL size (uc2-base 12 layers, XLM-R vocabulary 250 002, 120 pairs = 40 x
(1 positive + 2 negatives), up to 60 tokens + 100 regions) where the oracle is too slow to be the checker:
size-independent properties of the path instead.

  * batch-permutation equivariance of the scores (every kernel is row-local or per-sample),
  * padding invariance: a sample scores the same inside a long padded batch as inside a short one,
  * the triplet loss is the reference's formula of the scores (model/itm.py:43-53),
  * gradient sparsity is index-exact: only vocabulary rows of tokens in the batch (never padding_idx) and only the
    position rows in use receive gradient.
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_cfg2_full_size_properties():
    from uc2_b200 import batch as B, itm, synth
    from uc2_b200.config import UC2Config, retrieval_shapes
    from uc2_b200.utils import set_dropout
    cfg = UC2Config()
    assert cfg.num_hidden_layers == 12 and cfg.vocab_size == 250002
    m = itm.VLXLMRForImageTextRetrieval(cfg, 2048, margin=0.2)
    m.load_state_dict(synth.fill_state_dict(retrieval_shapes(cfg), seed=42, perturb=True), strict=False)
    m.cuda()
    set_dropout(m, 0)
    items = synth.make_pairs(120, seed=77)                     # ragged: 8..60 tokens, 10..100 regions
    batch = B.collate_itm_rank(items, 3)
    assert batch["attn_masks"].shape[0] == 120
    m.eval()
    with torch.no_grad():
        s = m(B.to_device(batch, "cuda"), compute_loss=False).float().reshape(-1)
        assert s.shape == (120,) and bool(torch.isfinite(s).all())
        assert float(s.std()) > 1e-3                            # not a constant: the comparison below means something
        # (1) permutation equivariance
        perm = torch.from_numpy(np.random.RandomState(5).permutation(120))
        s_perm = m(B.to_device(B.collate_itm_rank([items[i] for i in perm.tolist()], 3), "cuda"),
                   compute_loss=False).float().reshape(-1)
        assert float((s_perm - s[perm.cuda()]).abs().max()) <= 2e-3
        # (2) padding invariance: the 30 shortest samples alone (smaller T, R and S) vs inside the full batch
        order = sorted(range(120), key=lambda i: items[i]["input_ids"].numel() + items[i]["img_feat"].size(0))[:30]
        short = B.collate_itm_rank([items[i] for i in order], 3)
        assert short["attn_masks"].shape[1] < batch["attn_masks"].shape[1]
        s_short = m(B.to_device(short, "cuda"), compute_loss=False).float().reshape(-1)
        assert float((s_short - s[torch.tensor(order).cuda()]).abs().max()) <= 2e-2
    # (3) the loss is the triplet formula of the scores
    m.train()
    dev_batch = B.to_device(batch, "cuda")
    loss = m(dev_batch, compute_loss=True)
    assert loss.shape == (40, 2)
    with torch.no_grad():
        sg = torch.sigmoid(m(dev_batch, compute_loss=False).float()).view(-1, 3)
        want = torch.clamp(0.2 + sg[:, 1:] - sg[:, :1], 0)
    np.testing.assert_allclose(loss.detach().cpu().numpy(), want.cpu().numpy(), atol=2e-3)
    # (4) index-exact gradient sparsity
    loss.mean().backward()
    arena = m._arena()
    gw = arena.g("roberta.embeddings.word_embeddings.weight")
    used = torch.unique(batch["input_ids"])
    used = used[used != 1]                                      # padding_idx of the XLM-R table
    rows = gw.abs().sum(1).nonzero().squeeze(1).cpu()
    # a group whose two hinge terms are both clamped to 0 sends no gradient to its tokens, hence subset + coverage
    assert bool(torch.isin(rows, used).all()), "a vocabulary row outside the batch received gradient"
    assert rows.numel() >= used.numel() // 2, (rows.numel(), used.numel())
    gp = arena.g("roberta.embeddings.position_embeddings.weight")
    T = batch["input_ids"].shape[1]
    prow = gp.abs().sum(1).nonzero().squeeze(1).cpu()
    # XLM-R positions: padding_idx + 1 + index for real tokens (model/model.py:308-320); pad tokens sit at padding_idx
    assert int(prow.min()) >= 1 and int(prow.max()) == T + 1 and prow.numel() <= T + 1
    for n in ("roberta.encoder.layer.0.attention.self.query.weight", "roberta.encoder.layer.11.output.dense.weight",
              "roberta.img_embeddings.img_linear.weight", "rank_output.weight"):
        g = arena.g(n)
        assert bool(torch.isfinite(g).all()) and float(g.abs().max()) > 0, n
