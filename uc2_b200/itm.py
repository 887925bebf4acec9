"""Image-text retrieval models: mirrors of model/itm.py:12-55 (VLXLMR), 57-102 (Uniter) and the
in-batch hard-negative variant 105-186."""
from collections import defaultdict

import torch
from torch import nn

from . import functional as Fn
from .model import UC2PreTrainedModel, UniterModel, VLXLMRModel, _adopt


class _ForImageTextRetrieval(UC2PreTrainedModel):
    encoder_attr = "roberta"
    EncoderCls = VLXLMRModel

    def __init__(self, config, img_dim, margin=0.2):
        super().__init__(config)
        setattr(self, self.encoder_attr, self.EncoderCls(config, img_dim))
        self.itm_output = nn.Linear(config.hidden_size, 2)
        self.rank_output = nn.Linear(config.hidden_size, 1)
        self.margin = margin
        self.apply(self.init_weights)
        _adopt(self)

    @property
    def _enc(self):
        return getattr(self, self.encoder_attr)

    def init_output(self):
        """need to be called after from pretrained (model/itm.py:23-26): rank_output <- row 1 of itm_output.
        The reference rebinds .data to a slice; copying keeps the arena views intact with the same values."""
        with torch.no_grad():
            self.rank_output.weight.copy_(self.itm_output.weight[1:, :])
            self.rank_output.bias.copy_(self.itm_output.bias[1:])

    def forward(self, batch, compute_loss=True):
        batch = defaultdict(lambda: None, batch)
        position_ids = None if self.family_name == "vlxlmr" else batch["position_ids"]
        seq = self._enc(batch["input_ids"], position_ids, batch["img_feat"], batch["img_pos_feat"],
                        batch["attn_masks"], batch["gather_index"], output_all_encoded_layers=False)
        pooled = self._enc.pooler(seq)
        rank_scores = Fn.NarrowLinearFn.apply(pooled, self._arena(), "rank_output.weight", "rank_output.bias")
        if compute_loss:
            return Fn.RankLossFn.apply(rank_scores, int(batch["sample_size"]), float(self.margin))
        return rank_scores


class VLXLMRForImageTextRetrieval(_ForImageTextRetrieval):
    family_name = "vlxlmr"
    encoder_attr = "roberta"
    EncoderCls = VLXLMRModel


class UniterForImageTextRetrieval(_ForImageTextRetrieval):
    family_name = "uniter"
    encoder_attr = "bert"
    EncoderCls = UniterModel


class _HardNegMixin(object):
    """In-batch hard-negative mining (the behaviour of model/itm.py:105-186).  A batch holds ONE positive pair (row 0)
    and its candidates: one caption against many images (sample_from='t') or one image against many captions ('i').
    In training the candidates are first scored without gradient in eval mode, the `hard_size` highest-scoring
    negatives are kept next to the positive, and the triplet loss is taken on that reduced batch."""

    _FIXED = {"t": ("input_ids",), "i": ("img_feat", "img_pos_feat")}        # the side shared by every candidate

    def forward(self, batch, sample_from="t", compute_loss=True):
        if sample_from not in self._FIXED:
            raise ValueError()
        batch = dict(batch)
        n = batch["attn_masks"].size(0)
        for key in self._FIXED[sample_from]:                 # a single shared row is broadcast over the candidates
            if batch[key].size(0) == 1:
                batch[key] = batch[key].expand(n, *([-1] * (batch[key].dim() - 1)))
        if not (self.training and compute_loss):
            return super().forward(batch, compute_loss)
        with torch.no_grad():
            self.eval()
            try:
                scores = super().forward(batch, compute_loss=False)
                mined = self._get_hard_batch(batch, scores, sample_from)
            finally:
                self.train()
        return super().forward(mined, compute_loss=True)

    def _get_hard_batch(self, batch, scores, sample_from="t"):
        k = self.hard_size
        # row 0 is the positive; the k best-scoring other rows are the hard negatives
        keep = torch.cat([scores.new_zeros(1, dtype=torch.long),
                          scores.reshape(-1)[1:].topk(k, sorted=False).indices + 1])
        take = lambda t: t.index_select(0, keep)
        head = lambda t: t[:k + 1]
        attn, gidx = take(batch["attn_masks"]), take(batch["gather_index"])
        pos = batch.get("position_ids")
        if pos is not None and pos.size(0) != 1:
            pos = head(pos)
        if sample_from == "t":
            # images vary: trim the joint length to the longest kept pair (one host sync, as in the reference)
            ids = head(batch["input_ids"])
            longest = int(attn.sum(dim=1).max().item())
            n_img = longest - ids.size(1)
            attn, gidx = attn[:, :longest], gidx[:, :longest]
            feat, box = take(batch["img_feat"])[:, :n_img], take(batch["img_pos_feat"])[:, :n_img]
        else:
            ids = take(batch["input_ids"])
            feat, box = head(batch["img_feat"]), head(batch["img_pos_feat"])
        return {"sample_size": k + 1, "input_ids": ids, "position_ids": pos, "img_feat": feat, "img_pos_feat": box,
                "attn_masks": attn, "gather_index": gidx}


class UniterForImageTextRetrievalHardNeg(_HardNegMixin, UniterForImageTextRetrieval):
    def __init__(self, config, img_dim, margin=0.2, hard_size=16):
        super().__init__(config, img_dim, margin)
        self.hard_size = hard_size


class VLXLMRForImageTextRetrievalHardNeg(_HardNegMixin, VLXLMRForImageTextRetrieval):
    """The reference ships the hard-negative variant for the Uniter family only; the same in-batch mining on
    the VLXLMR model is what BASELINE.json configs[1] ("ITM with hard negatives") exercises."""

    def __init__(self, config, img_dim, margin=0.2, hard_size=16):
        super().__init__(config, img_dim, margin)
        self.hard_size = hard_size
