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
    """model/itm.py:105-186: score all candidates without grad, keep the hard_size highest-scoring
    negatives plus the positive (row 0), train on those."""

    def forward(self, batch, sample_from="t", compute_loss=True):
        batch = dict(batch)
        batch_size = batch["attn_masks"].size(0)
        input_ids, img_feat, img_pos_feat = batch["input_ids"], batch["img_feat"], batch["img_pos_feat"]
        if sample_from == "t":
            if input_ids.size(0) == 1:
                batch["input_ids"] = input_ids.expand(batch_size, -1)
        elif sample_from == "i":
            if img_feat.size(0) == 1:
                batch["img_feat"] = img_feat.expand(batch_size, -1, -1)
            if img_pos_feat.size(0) == 1:
                batch["img_pos_feat"] = img_pos_feat.expand(batch_size, -1, -1)
        else:
            raise ValueError()
        if self.training and compute_loss:
            with torch.no_grad():
                self.eval()
                scores = super().forward(batch, compute_loss=False)
                hard_batch = self._get_hard_batch(batch, scores, sample_from)
                self.train()
            return super().forward(hard_batch, compute_loss=True)
        return super().forward(batch, compute_loss)

    def _get_hard_batch(self, batch, scores, sample_from="t"):
        batch = defaultdict(lambda: None, batch)
        input_ids, position_ids = batch["input_ids"], batch["position_ids"]
        img_feat, img_pos_feat = batch["img_feat"], batch["img_pos_feat"]
        attention_mask, gather_index = batch["attn_masks"], batch["gather_index"]
        hard_batch = {"sample_size": self.hard_size + 1}
        # first example is the positive
        hard_indices = scores.squeeze(-1)[1:].topk(self.hard_size, sorted=False)[1] + 1
        indices = torch.cat([torch.zeros(1, dtype=torch.long, device=hard_indices.device), hard_indices])
        attention_mask = attention_mask.index_select(0, indices)
        gather_index = gather_index.index_select(0, indices)
        if position_ids is not None and position_ids.size(0) != 1:
            position_ids = position_ids[:self.hard_size + 1]
        if sample_from == "t":
            max_len = int(attention_mask.sum(dim=1).max().item())     # cut to minimum padding
            max_i = max_len - input_ids.size(1)
            attention_mask = attention_mask[:, :max_len]
            gather_index = gather_index[:, :max_len]
            img_feat = img_feat.index_select(0, indices)[:, :max_i, :]
            img_pos_feat = img_pos_feat.index_select(0, indices)[:, :max_i, :]
            input_ids = input_ids[:self.hard_size + 1]
        elif sample_from == "i":
            input_ids = input_ids.index_select(0, indices)
            img_feat = img_feat[:self.hard_size + 1]
            img_pos_feat = img_pos_feat[:self.hard_size + 1]
        else:
            raise ValueError()
        hard_batch.update(input_ids=input_ids, position_ids=position_ids, img_feat=img_feat,
                          img_pos_feat=img_pos_feat, attn_masks=attention_mask, gather_index=gather_index)
        return hard_batch


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
