"""Host-side batch assembly for the UC2 encoder path (vectorised restatement of the
reference collate index builders; the per-sample Python loops of the reference
become closed-form index arithmetic).

Reference behaviour followed (file:line under /root/reference):
  pad_tensors            data/data.py:360-373
  get_gather_index       data/data.py:376-384
  _compute_ot_scatter    data/itm.py:264-271
  _compute_pad           data/itm.py:274-278
  xlmr_itm_ot_collate    data/itm.py:281-319
  xlmr_itm_rank_collate  data/itm.py:615-643
  mrfr / mrc collates    data/mrm.py:22-39, 213-218
  mlm collate            data/mlm.py (txt_labels padded with -1)
"""
import torch
from torch.nn.utils.rnn import pad_sequence


def pad_tensors(tensors, lens=None, pad=0):
    """B x [n_i, D] -> [B, max n_i, D], zero (or ``pad``) filled."""
    if lens is None:
        lens = [t.size(0) for t in tensors]
    out = tensors[0].new_full((len(tensors), max(lens), tensors[0].size(-1)), pad)
    for i, (t, l) in enumerate(zip(tensors, lens)):
        out[i, :l] = t[:l]
    return out


def get_gather_index(txt_lens, num_bbs, batch_size, max_len, out_size):
    """gather_index[b, j] = j, except j in [tl_b, tl_b+nbb_b) -> max_len + (j - tl_b).

    Pad columns (j >= tl_b + nbb_b) keep the identity index, exactly like the
    reference."""
    assert len(txt_lens) == len(num_bbs) == batch_size
    tl = torch.as_tensor(txt_lens, dtype=torch.long).unsqueeze(1)
    nb = torch.as_tensor(num_bbs, dtype=torch.long).unsqueeze(1)
    j = torch.arange(out_size, dtype=torch.long).unsqueeze(0).expand(batch_size, -1)
    in_img = (j >= tl) & (j < tl + nb)
    return torch.where(in_img, j - tl + max_len, j).contiguous()


def compute_ot_scatter(txt_lens, max_txt_len, joint_len):
    """ot_scatter[b, j] = j for j < tl_b else max_txt_len + (j - tl_b)."""
    tl = torch.as_tensor(txt_lens, dtype=torch.long).unsqueeze(1)
    j = torch.arange(joint_len, dtype=torch.long).unsqueeze(0).expand(len(txt_lens), -1)
    return torch.where(j >= tl, j - tl + max_txt_len, j).contiguous()


def compute_pad(lens, max_len, dtype=torch.bool):
    """pad[b, j] = (j >= len_b). The reference emits uint8 (torch<1.2 idiom);
    bool is the same mask and is what torch>=2 masked_fill_ accepts."""
    l = torch.as_tensor(lens, dtype=torch.long).unsqueeze(1)
    return (torch.arange(max_len).unsqueeze(0) >= l).to(dtype)


def _common(items, pad_id=1):
    input_ids = [it["input_ids"] for it in items]
    txt_lens = [int(i.numel()) for i in input_ids]
    num_bbs = [int(it["img_feat"].size(0)) for it in items]
    ids = pad_sequence(input_ids, batch_first=True, padding_value=pad_id)
    position_ids = torch.arange(0, ids.size(1), dtype=torch.long).unsqueeze(0)
    img_feat = pad_tensors([it["img_feat"] for it in items], num_bbs)
    img_pos_feat = pad_tensors([it["img_pos_feat"] for it in items], num_bbs)
    attn = pad_sequence([torch.ones(t + n, dtype=torch.long) for t, n in zip(txt_lens, num_bbs)],
                        batch_first=True, padding_value=0)
    bs, max_tl = ids.shape
    gather_index = get_gather_index(txt_lens, num_bbs, bs, max_tl, attn.size(1))
    batch = dict(input_ids=ids, position_ids=position_ids, img_feat=img_feat,
                 img_pos_feat=img_pos_feat, attn_masks=attn, gather_index=gather_index)
    return batch, txt_lens, num_bbs


def collate_itm(items, targets, with_ot=True, pad_id=1):
    batch, txt_lens, num_bbs = _common(items, pad_id)
    batch["targets"] = torch.as_tensor(targets, dtype=torch.long)
    if with_ot:
        max_tl, max_nbb = max(txt_lens), max(num_bbs)
        ot_scatter = compute_ot_scatter(txt_lens, max_tl, batch["attn_masks"].size(1))
        batch["ot_inputs"] = dict(ot_scatter=ot_scatter, scatter_max=int(ot_scatter.max()),
                                  txt_pad=compute_pad(txt_lens, max_tl),
                                  img_pad=compute_pad(num_bbs, max_nbb))
    return batch


def collate_itm_rank(items, sample_size, pad_id=1):
    """Triplet-ranking batch: consecutive groups of ``sample_size`` pairs, first is positive."""
    batch, _, _ = _common(items, pad_id)
    assert len(items) % sample_size == 0
    batch["sample_size"] = sample_size
    return batch


def collate_mlm(items, labeled, pad_id=1):
    """``labeled``: list of (masked_ids, labels) aligned with items."""
    its = [dict(it, input_ids=m) for it, (m, _) in zip(items, labeled)]
    batch, _, _ = _common(its, pad_id)
    batch["txt_labels"] = pad_sequence([l for _, l in labeled], batch_first=True, padding_value=-1)
    return batch


def tlm_position_ids(input_ids, start=2):
    """data/mlm.py:420-428 (VTLM): positions count up from 3 and restart at every <s> (id 0), so both halves of a
    translation pair get the same position range."""
    out, pos = [], start
    for t in input_ids.tolist():
        pos = start if t == 0 else pos + 1
        out.append(pos)
    return torch.tensor(out, dtype=torch.long)


def collate_tlm(items, labeled, pad_id=1):
    """xlmr_mlm_dmasking_collate (data/mlm.py:845-884): as collate_mlm plus per-sample position ids padded with 1.
    (The text-only sibling xlmr_tlm_ni_dmasking_collate, 803-842, differs only in gather_index = None; run it as task
    'tlm-ni', which ignores the image side.)"""
    batch = collate_mlm(items, labeled, pad_id)
    batch["position_ids"] = pad_sequence([tlm_position_ids(m) for m, _ in labeled], batch_first=True, padding_value=1)
    return batch


def _mrm_common(items, img_masks, pad_id=1):
    batch, txt_lens, num_bbs = _common(items, pad_id)
    m = pad_sequence(img_masks, batch_first=True, padding_value=0).bool()
    # img_mask_tgt lives in packed coordinates: zeros(tl) ++ mask (data/mrm.py:22-25)
    tgt = pad_sequence([torch.cat([torch.zeros(t, dtype=torch.bool), mk])
                        for t, mk in zip(txt_lens, img_masks)], batch_first=True, padding_value=0)
    S = batch["attn_masks"].size(1)
    if tgt.size(1) < S:
        tgt = torch.cat([tgt, tgt.new_zeros(tgt.size(0), S - tgt.size(1))], 1)
    batch["img_masks"], batch["img_mask_tgt"] = m, tgt
    return batch, m


def collate_mrfr(items, img_masks, pad_id=1):
    batch, m = _mrm_common(items, img_masks, pad_id)
    feat = batch["img_feat"]
    batch["feat_targets"] = feat[m].contiguous()          # row-major (b, r) order
    batch["img_feat"] = feat.masked_fill(m.unsqueeze(-1), 0)
    return batch


def collate_mrc(items, img_masks, soft_labels, pad_id=1):
    batch, m = _mrm_common(items, img_masks, pad_id)
    lab = pad_tensors(soft_labels, [s.size(0) for s in soft_labels])
    batch["label_targets"] = lab[m].contiguous()
    batch["img_feat"] = batch["img_feat"].masked_fill(m.unsqueeze(-1), 0)
    return batch


def collate_mmxlm(items, labeled, img_masks, img_token_labels, pad_id=1):
    """xlmr_mmxlm_collate (data/mlm.py:887-934), tasks 'mmxlm' / 'vmlm' (MRTM): txt_labels live in PACKED coordinates,
    cat(caption labels [tl], region token labels [nbb]) per sample (data/mlm.py:467), padded with -1; masked regions
    are zeroed (_mask_img_feat) and flagged in img_masks."""
    its = [dict(it, input_ids=m) for it, (m, _) in zip(items, labeled)]
    batch, m = _mrm_common(its, img_masks, pad_id)
    del batch["img_mask_tgt"]
    batch["txt_labels"] = pad_sequence([torch.cat([l, r]) for (_, l), r in zip(labeled, img_token_labels)],
                                       batch_first=True, padding_value=-1)
    S = batch["attn_masks"].size(1)
    if batch["txt_labels"].size(1) < S:
        batch["txt_labels"] = torch.nn.functional.pad(batch["txt_labels"], (0, S - batch["txt_labels"].size(1)), value=-1)
    batch["img_feat"] = batch["img_feat"].masked_fill(m.unsqueeze(-1), 0)
    batch["n_masked"] = int((batch["txt_labels"] != -1).sum())
    return batch


def collate_mmxlm_soft(items, img_masks, img_token_soft_labels, pad_id=1):
    """xlmr_mmxlm_softlabel_collate (data/mlm.py:936-1008), tasks 'mmxlm-soft' / 'vmlm-soft': tgt_masks marks the
    masked regions in packed coordinates, label_targets = rows of the per-region token distributions at the masked
    regions (_get_targets), row-major."""
    batch, m = _mrm_common(items, img_masks, pad_id)
    batch["tgt_masks"] = batch.pop("img_mask_tgt")
    soft = pad_tensors(img_token_soft_labels, [s.size(0) for s in img_token_soft_labels])
    batch["label_targets"] = soft[m].contiguous()
    batch["img_feat"] = batch["img_feat"].masked_fill(m.unsqueeze(-1), 0)
    return batch


def to_device(batch, device, non_blocking=True):
    out = {}
    for k, v in batch.items():
        if torch.is_tensor(v):
            out[k] = v.to(device, non_blocking=non_blocking)
        elif isinstance(v, dict):
            out[k] = to_device(v, device, non_blocking)
        else:
            out[k] = v
    return out


class Prefetcher(object):
    """data/loader.py:75-135 (PrefetchLoader): copies the NEXT batch host->device on a side stream while the
    current step computes.  Batches must be pinned for the copy to overlap.

    Unlike the reference, the device buffers are allocated on the COMPUTE stream (the side stream first waits for
    whatever that stream had queued, i.e. the previous step) and only the copies run on the side stream: memory
    allocated on one stream and consumed on another cannot be reused by the caching allocator until events
    resolve, and with four alternating task batches that turned into a cudaMalloc -- a device-wide sync -- per
    step, serialising the copy it was supposed to hide."""

    def __init__(self, loader, device):
        self.loader = loader
        self.device = torch.device(device)
        self.stream = torch.cuda.Stream(device=self.device)

    def __len__(self):
        return len(self.loader)

    def _preload(self, it):
        try:
            host = next(it)
        except StopIteration:
            return None
        cur = torch.cuda.current_stream(self.device)
        dst = self._alloc(host)                      # on the compute stream's pool
        self.stream.wait_stream(cur)                 # the buffers may still be in use by already queued work
        with torch.cuda.stream(self.stream):
            self._copy(dst, host)
        return dst

    def _alloc(self, b):
        if torch.is_tensor(b):
            return torch.empty(b.shape, dtype=b.dtype, device=self.device)
        if isinstance(b, dict):
            return {k: self._alloc(v) for k, v in b.items()}
        if isinstance(b, (list, tuple)):
            return type(b)(self._alloc(v) for v in b)
        return b

    def _copy(self, d, h):
        if torch.is_tensor(h):
            d.copy_(h, non_blocking=True)
        elif isinstance(h, dict):
            for k in h:
                self._copy(d[k], h[k])
        elif isinstance(h, (list, tuple)):
            for dv, hv in zip(d, h):
                self._copy(dv, hv)

    def __iter__(self):
        it = iter(self.loader)
        nxt = self._preload(it)
        while nxt is not None:
            torch.cuda.current_stream(self.device).wait_stream(self.stream)     # batch i is on the device
            batch = nxt
            nxt = self._preload(it)                  # batch i+1: copied while step i computes
            yield batch
