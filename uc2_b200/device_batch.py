"""Device-side batch assembly: the collate step before the encoder path, for region features
that live in HBM.

The reference reads each image's features from LMDB, pads them per sample on the host (pad_tensors,
data/data.py:360-373) and ships B x R x 8 KB of fp32 through pinned memory every step; its index tensors come from
per-sample Python loops (get_gather_index data/data.py:376-384, _compute_ot_scatter / _compute_pad
data/itm.py:264-278, _get_img_tgt_mask / _get_feat_target / _mask_img_feat data/mrm.py:22-39, _get_targets 213-218).
A B200 has 180 GB: COCO + VG region features (≈ 50 GB fp32) fit, so `FeatureArena` keeps the whole store resident
as one ragged matrix and `DeviceCollator` turns (token ids, image indices, region masks) -- a few KB of host data per
step -- into exactly the batch dict `uc2_b200.batch.collate_*` + `to_device` would have produced (bit-identical:
copies and integer arithmetic only; checked in tests/test_device_batch_gpu.py).
"""
import torch
from torch.nn.utils.rnn import pad_sequence

from . import _lib
from ._lib import call, ptr, stream
from .batch import tlm_position_ids


class FeatureArena(object):
    """Ragged HBM-resident feature store: image i owns rows [row0[i], row0[i] + nbb[i]) of
    `feat` [N, 2048], `pos` [N, 7] and, when given, `soft` [N, C] (the MRC soft labels)."""

    def __init__(self, feats, boxes, soft_labels=None, device="cuda"):
        assert len(feats) == len(boxes) and len(feats) > 0
        self.nbb = [int(f.size(0)) for f in feats]
        row0 = [0]
        for n in self.nbb[:-1]:
            row0.append(row0[-1] + n)
        self.row0 = row0
        self.device = torch.device(device)
        self.feat = torch.cat([f.float() for f in feats]).contiguous().to(self.device)
        self.pos = torch.cat([b.float() for b in boxes]).contiguous().to(self.device)
        self.soft = None
        if soft_labels is not None:
            assert [int(s.size(0)) for s in soft_labels] == self.nbb
            self.soft = torch.cat([s.float() for s in soft_labels]).contiguous().to(self.device)

    def __len__(self):
        return len(self.nbb)

    @classmethod
    def synthetic(cls, n_images, bb_range=(10, 100), seed=0, device="cuda", img_dim=2048, soft_dim=0):
        """Generated on the device (bench only): |N(0,1)| features, sorted-uniform boxes (data/data.py:339)."""
        self = cls.__new__(cls)
        g = torch.Generator().manual_seed(seed)
        nbb = torch.randint(bb_range[0], bb_range[1] + 1, (n_images,), generator=g)
        self.nbb = nbb.tolist()
        self.row0 = (torch.cumsum(nbb, 0) - nbb).tolist()
        self.device = torch.device(device)
        n = int(nbb.sum())
        gd = torch.Generator(device=device).manual_seed(seed)
        self.feat = torch.randn((n, img_dim), generator=gd, device=device).abs_()
        c = torch.rand((n, 4), generator=gd, device=device)
        x1, x2 = torch.minimum(c[:, 0], c[:, 1]), torch.maximum(c[:, 0], c[:, 1])
        y1, y2 = torch.minimum(c[:, 2], c[:, 3]), torch.maximum(c[:, 2], c[:, 3])
        self.pos = torch.stack([x1, y1, x2, y2, x2 - x1, y2 - y1, (x2 - x1) * (y2 - y1)], -1).contiguous()
        self.soft = None
        if soft_dim:
            self.soft = torch.softmax(torch.randn((n, soft_dim), generator=gd, device=device) * 3, -1).contiguous()
        return self


class DeviceCollator(object):
    """Builds the reference's batch dicts on the device from an arena.  Every method takes per-sample token id
    tensors (host, as the text DB yields them) and `img_idx`, the arena index of each sample's image."""

    def __init__(self, arena, pad_id=1):
        self.arena = arena
        self.pad_id = pad_id

    # ------------------------------------------------------------------ internals
    def _pad(self, src, row0, nbb, B, R, mask=None, slot=None, n_tgt=0, zero_masked=False, want_out=True):
        D = src.size(1)
        dev = src.device
        out = torch.empty((B, R, D), dtype=torch.float32, device=dev) if want_out else None
        tgt = torch.empty((n_tgt, D), dtype=torch.float32, device=dev) if n_tgt else None
        a = _lib.PadArgs(arena=ptr(src), D=D, row0=ptr(row0), nbb=ptr(nbb), mask=ptr(mask),
                         tgt_slot=ptr(slot) if n_tgt else None, B=B, R=R, zero_masked=int(zero_masked))
        call("uc2_pad_rows", a, ptr(out), ptr(tgt), stream())
        return out, tgt

    def _assemble(self, input_ids, img_idx, img_masks=None, targets=None, zero_masked=False, ot=False,
                  tgt_mask=False, position_ids=None):
        ar, dev = self.arena, self.arena.device
        B = len(input_ids)
        assert len(img_idx) == B
        tl = [int(t.numel()) for t in input_ids]
        nbb = [ar.nbb[i] for i in img_idx]
        T, R = max(tl), max(nbb)
        S = max(t + n for t, n in zip(tl, nbb))
        # everything the device needs to know about this batch: 3 integers per sample + the token ids
        meta = torch.tensor([[ar.row0[i] for i in img_idx], nbb, tl], dtype=torch.int64).to(dev, non_blocking=True)
        row0 = meta[0].contiguous()
        nbb_d, tl_d = meta[1].int(), meta[2].int()
        ids = pad_sequence(list(input_ids), batch_first=True, padding_value=self.pad_id).to(dev, non_blocking=True)
        if position_ids is None:
            position_ids = torch.arange(0, T, dtype=torch.long, device=dev).unsqueeze(0)
        mask_d = slot = None
        n_masked = 0
        if img_masks is not None:
            m = pad_sequence(list(img_masks), batch_first=True, padding_value=0).bool()
            if m.size(1) < R:
                m = torch.nn.functional.pad(m, (0, R - m.size(1)))
            n_masked = int(m.sum())
            mask_d = m.to(torch.uint8).to(dev, non_blocking=True).contiguous()
            slot = (torch.cumsum(mask_d.flatten().int(), 0) - 1).int()       # exclusive row-major scan
        want_feat_tgt = targets == "feat"
        feat, feat_tgt = self._pad(ar.feat, row0, nbb_d, B, R, mask_d, slot, n_masked if want_feat_tgt else 0,
                                   zero_masked)
        pos, _ = self._pad(ar.pos, row0, nbb_d, B, R)
        attn = torch.empty((B, S), dtype=torch.long, device=dev)
        gi = torch.empty((B, S), dtype=torch.long, device=dev)
        ot_sc = torch.empty((B, S), dtype=torch.long, device=dev) if ot else None
        tpad = torch.empty((B, T), dtype=torch.uint8, device=dev) if ot else None
        ipad = torch.empty((B, R), dtype=torch.uint8, device=dev) if ot else None
        mt = torch.empty((B, S), dtype=torch.uint8, device=dev) if tgt_mask else None
        call("uc2_batch_index", ptr(tl_d), ptr(nbb_d), ptr(mask_d), B, T, R, S, ptr(attn), ptr(gi), ptr(ot_sc),
             ptr(tpad), ptr(ipad), ptr(mt), stream())
        batch = dict(input_ids=ids, position_ids=position_ids, img_feat=feat, img_pos_feat=pos, attn_masks=attn,
                     gather_index=gi)
        if ot:
            batch["ot_inputs"] = dict(ot_scatter=ot_sc, scatter_max=S - 1 - min(tl) + T,
                                      txt_pad=tpad.view(torch.bool), img_pad=ipad.view(torch.bool))
        if img_masks is not None:
            batch["img_masks"] = mask_d.view(torch.bool)
        if tgt_mask:
            batch["img_mask_tgt"] = mt.view(torch.bool)
        if want_feat_tgt:
            batch["feat_targets"] = feat_tgt if feat_tgt is not None else feat.new_zeros((0, feat.size(-1)))
        if targets == "soft":
            assert ar.soft is not None, "the arena holds no soft labels"
            _, lab = self._pad(ar.soft, row0, nbb_d, B, R, mask_d, slot, n_masked, want_out=False)
            batch["label_targets"] = lab if lab is not None else feat.new_zeros((0, ar.soft.size(-1)))
        return batch, tl, nbb

    # ------------------------------------------------------------------ the reference's collates
    def itm(self, input_ids, img_idx, targets, with_ot=True):
        """xlmr_itm_ot_collate data/itm.py:281-319."""
        batch, _, _ = self._assemble(input_ids, img_idx, ot=with_ot)
        batch["targets"] = torch.as_tensor(targets, dtype=torch.long).to(self.arena.device, non_blocking=True)
        return batch

    def itm_rank(self, input_ids, img_idx, sample_size):
        """xlmr_itm_rank_collate data/itm.py:615-643."""
        assert len(input_ids) % sample_size == 0
        batch, _, _ = self._assemble(input_ids, img_idx)
        batch["sample_size"] = sample_size
        return batch

    def _labels(self, labels, S=None):
        lab = pad_sequence(list(labels), batch_first=True, padding_value=-1)
        if S is not None and lab.size(1) < S:
            lab = torch.nn.functional.pad(lab, (0, S - lab.size(1)), value=-1)
        return lab.to(self.arena.device, non_blocking=True)

    def mlm(self, labeled, img_idx):
        """xlmr_mlm_collate (data/mlm.py): `labeled` = [(masked_ids, labels)], txt_labels padded with -1."""
        batch, _, _ = self._assemble([m for m, _ in labeled], img_idx)
        batch["txt_labels"] = self._labels([l for _, l in labeled])
        return batch

    def tlm(self, labeled, img_idx):
        """xlmr_tlm_ni_dmasking_collate data/mlm.py:803-842: per-sample position ids padded with 1."""
        pos = pad_sequence([tlm_position_ids(m) for m, _ in labeled], batch_first=True, padding_value=1)
        batch, _, _ = self._assemble([m for m, _ in labeled], img_idx,
                                     position_ids=pos.to(self.arena.device, non_blocking=True))
        batch["txt_labels"] = self._labels([l for _, l in labeled])
        return batch

    def mrfr(self, input_ids, img_idx, img_masks):
        """mrfr_collate data/mrm.py:88-123: feat_targets = the masked rows before zeroing, row-major."""
        batch, _, _ = self._assemble(input_ids, img_idx, img_masks, targets="feat", zero_masked=True, tgt_mask=True)
        return batch

    def mrc(self, input_ids, img_idx, img_masks):
        """mrc_collate data/mrm.py:258-297: label_targets = soft labels of the masked regions."""
        batch, _, _ = self._assemble(input_ids, img_idx, img_masks, targets="soft", zero_masked=True, tgt_mask=True)
        return batch

    def mmxlm(self, labeled, img_idx, img_masks, img_token_labels):
        """xlmr_mmxlm_collate data/mlm.py:887-934 (tasks mmxlm / vmlm): txt_labels in packed coordinates."""
        batch, tl, nbb = self._assemble([m for m, _ in labeled], img_idx, img_masks, zero_masked=True)
        S = batch["attn_masks"].size(1)
        lab = [torch.cat([l, r]) for (_, l), r in zip(labeled, img_token_labels)]
        batch["txt_labels"] = self._labels(lab, S)
        batch["n_masked"] = int(sum(int((x != -1).sum()) for x in lab))
        return batch
