"""Per-sample random draws that sit between the datasets and the collates: BERT-style token masking, region masking
and negative sampling.  They consume Python's `random` stream in the same order as the reference, so a run seeded like
the reference (utils/misc.py set_random_seed) draws the same masks and the same negatives; tests/golden/sampling.npz
(produced by the reference's own functions) pins that.

  random_word        data/mlm.py:30-67
  get_img_mask       data/mrm.py:13-19
  sample_negative    data/itm.py:39-44 (the `_sample_negative_rand` implementation the reference selects at line 58)
  rank_id_pairs      data/itm.py:380-393 (ItmRankDataset.__getitem__: 1 positive + n image negatives + n text negatives)
"""
import random

import torch


def random_word(tokens, vocab_range, mask):
    """15 % of the tokens are selected; of those 80 % become `mask`, 10 % a uniformly drawn id of `vocab_range`,
    10 % stay.  Returns (tokens edited in place, labels with -1 where nothing is predicted); if nothing was selected
    the first token is masked.  The replacement id is drawn with random.choice over a range object, which consumes
    the generator exactly like the reference's choice over the materialised 250 k-element list."""
    lo_hi = range(*vocab_range)
    labels = [-1] * len(tokens)
    for i, tok in enumerate(tokens):
        p = random.random()
        if p >= 0.15:
            continue
        p /= 0.15
        if p < 0.8:
            tokens[i] = mask
        elif p < 0.9:
            tokens[i] = random.choice(lo_hi)
        labels[i] = tok
    if tokens and all(l == -1 for l in labels):
        labels[0], tokens[0] = tokens[0], mask
    return tokens, labels


def create_mlm_io(input_ids, vocab_range, mask, cls_, sep):
    """MlmDataset.create_mlm_io data/mlm.py:470-478: mask, then wrap in <s> ... </s> with unpredicted specials."""
    ids, labels = random_word(list(input_ids), vocab_range, mask)
    return torch.tensor([cls_] + ids + [sep]), torch.tensor([-1] + labels + [-1])


def get_img_mask(mask_prob, num_bb):
    """Bernoulli(mask_prob) per region, at least one region masked; bool tensor [num_bb]."""
    picks = [random.random() < mask_prob for _ in range(num_bb)]
    if not any(picks):
        picks[random.choice(range(num_bb))] = True
    return torch.tensor(picks)


def sample_negative(sample_pool, ground_truths, num_sample):
    """Draw `num_sample` items of the pool, retrying until none of them is a ground truth."""
    banned = set(ground_truths)
    while True:
        out = random.sample(sample_pool, num_sample)
        if banned.isdisjoint(out):
            return out


def rank_id_pairs(gt_txt_id, gt_img, img_pool, txt_pool, gt_img_txts, neg_sample_size=1):
    """(txt, img) id pairs of one retrieval training item: the positive, then `neg_sample_size` pairs with a wrong
    image, then `neg_sample_size` pairs with a wrong caption (a caption of another image)."""
    neg_imgs = sample_negative(img_pool, [gt_img], neg_sample_size)
    neg_txts = sample_negative(txt_pool, gt_img_txts, neg_sample_size)
    return [(gt_txt_id, gt_img)] + [(gt_txt_id, i) for i in neg_imgs] + [(t, gt_img) for t in neg_txts]
