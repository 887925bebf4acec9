"""Image-text retrieval scoring and evaluation on the B200 path.

Mirrors the reference's evaluation flow (file:line under /root/reference):

  ItmEvalDataset          data/itm.py:891-902   all images sorted by number of boxes, scored in chunks of
                                                `mini_batch_size` (400, config/uc2_mscoco_itm.json inf_minibatch_size)
  ItmValDataset.get_batch data/itm.py:456-485   one caption x a chunk of images -> one padded batch
  inference               itm.py:515-538        score_matrix[i, j:j+bs] = model(batch).squeeze(1).half()
  evaluate                itm.py:492-512        caption rows are sharded `ids[rank::world]`, rows all-gathered
  itm_eval                eval/itm.py:6-53      recall@1/5/10 in both directions

What changes (B200-first): the reference rebuilds every padded image chunk on the host for EVERY caption
(25 000 x 13 pad_tensors calls + 8 KB/region H2D each time).  Here the image side of every chunk is assembled once
into an HBM-resident `ImageArena` (5 000 images x <= 100 regions x 2048 fp32 = 2.3 GB of 180 GB), the caption
side (input_ids expand, attention mask, gather_index) is built on the device from (tl, nbb) vectors and cached per
(tl, chunk), and recall@k runs as vectorised device ops.  The score values, their fp16 storage, the row sharding and
the ranking rules are the reference's.
"""
import torch

from . import distributed as D


class ImageArena(object):
    """Device-resident image chunks.  `images`: list of dicts(img_feat [nbb, 2048], img_pos_feat [nbb, 7]) with
    optional 'id'.  Images are ordered by number of boxes (stable), as ItmEvalDataset does, and cut into chunks of
    `mini_batch_size`; each chunk is zero-padded to its own widest image (pad_tensors, data/data.py:360-373)."""

    def __init__(self, images, mini_batch_size=400, device="cuda", ids=None):
        nbb = [int(im["img_feat"].size(0)) for im in images]
        order = sorted(range(len(images)), key=lambda i: nbb[i])
        self.order = order
        self.img_ids = [ids[i] if ids is not None else images[i].get("id", i) for i in order]
        self.num_images = len(images)
        self.bs = int(mini_batch_size)
        self.chunks = []
        for st in range(0, len(order), self.bs):
            sel = order[st:st + self.bs]
            nb = [nbb[i] for i in sel]
            R = max(nb)
            feat = torch.zeros((len(sel), R, images[sel[0]]["img_feat"].size(-1)), dtype=torch.float32)
            pos = torch.zeros((len(sel), R, 7), dtype=torch.float32)
            for k, i in enumerate(sel):
                feat[k, :nb[k]] = images[i]["img_feat"]
                pos[k, :nb[k]] = images[i]["img_pos_feat"]
            self.chunks.append(dict(img_feat=feat.to(device), img_pos_feat=pos.to(device),
                                    num_bbs=torch.tensor(nb, dtype=torch.long, device=device), R=R, n=len(sel)))
        self._cache = {}

    @classmethod
    def synthetic(cls, n_images, mini_batch_size=400, bb_range=(10, 100), seed=0, device="cuda", img_dim=2048):
        """Synthetic evaluation set generated on the device (bench only): box counts ~ U{bb_range}, features
        |N(0,1)|, boxes sorted-uniform corners -> [x1,y1,x2,y2,w,h,w*h] (data/data.py:339)."""
        self = cls.__new__(cls)
        g = torch.Generator(device=device).manual_seed(seed)
        nbb = torch.randint(bb_range[0], bb_range[1] + 1, (n_images,), generator=g, device=device)
        nbb, order = torch.sort(nbb, stable=True)
        self.order = order.tolist()
        self.img_ids = list(self.order)
        self.num_images, self.bs, self.chunks, self._cache = n_images, int(mini_batch_size), [], {}
        for st in range(0, n_images, self.bs):
            nb = nbb[st:st + self.bs]
            n, R = nb.numel(), int(nb.max())
            valid = (torch.arange(R, device=device)[None, :] < nb[:, None]).unsqueeze(-1)
            feat = torch.randn((n, R, img_dim), generator=g, device=device).abs_() * valid
            c = torch.rand((n, R, 4), generator=g, device=device)
            x1, x2 = torch.minimum(c[..., 0], c[..., 1]), torch.maximum(c[..., 0], c[..., 1])
            y1, y2 = torch.minimum(c[..., 2], c[..., 3]), torch.maximum(c[..., 2], c[..., 3])
            pos = torch.stack([x1, y1, x2, y2, x2 - x1, y2 - y1, (x2 - x1) * (y2 - y1)], -1) * valid
            self.chunks.append(dict(img_feat=feat.contiguous(), img_pos_feat=pos.contiguous(), num_bbs=nb, R=R, n=n))
        return self

    def caption_side(self, c, tl):
        """attn_masks [n, tl+R] and gather_index [n, tl+R] of chunk c for a caption of tl tokens
        (data/itm.py:471-477), built on the device and cached."""
        key = (c, tl)
        hit = self._cache.get(key)
        if hit is None:
            ch = self.chunks[c]
            S = tl + ch["R"]
            j = torch.arange(S, device=ch["num_bbs"].device).unsqueeze(0)
            nb = ch["num_bbs"].unsqueeze(1)
            attn = (j < tl + nb).long()
            # get_gather_index([tl]*n, num_bbs, n, tl, S): every caption row has T == tl tokens, so the image block
            # [tl, tl+nbb) maps to T + (j - tl) = j and the index is the identity on every row
            gi = j.expand(ch["n"], -1).contiguous()
            hit = (attn.contiguous(), gi)
            if len(self._cache) > 4096:
                self._cache.clear()
            self._cache[key] = hit
        return hit

    def batch(self, c, input_ids):
        """The batch ItmValDataset.get_batch would build for (caption, chunk c); input_ids: 1-D device tensor."""
        ch = self.chunks[c]
        tl = int(input_ids.numel())
        attn, gi = self.caption_side(c, tl)
        return dict(input_ids=input_ids.unsqueeze(0).expand(ch["n"], -1).contiguous(),
                    position_ids=torch.arange(0, tl, dtype=torch.long, device=input_ids.device).unsqueeze(0),
                    img_feat=ch["img_feat"], img_pos_feat=ch["img_pos_feat"], attn_masks=attn, gather_index=gi)


@torch.no_grad()
def inference(model, captions, arena, rank=None, world=None):
    """itm.py:515-538.  `captions`: list of 1-D int64 tensors (token ids incl. <s> </s>).  Scores this rank's caption
    rows `captions[rank::world]` against every image; returns fp16 [n_local, n_images] in arena (sorted) order."""
    rank = D.rank() if rank is None else rank
    world = D.size() if world is None else world
    was_training = model.training
    model.eval()
    mine = list(range(rank, len(captions), world))
    dev = arena.chunks[0]["img_feat"].device
    score_matrix = torch.zeros((len(mine), arena.num_images), device=dev, dtype=torch.float16)
    for i, ci in enumerate(mine):
        ids = captions[ci].to(dev, non_blocking=True)
        j = 0
        for c in range(len(arena.chunks)):
            scores = model(arena.batch(c, ids), compute_loss=False)
            bs = scores.size(0)
            score_matrix[i, j:j + bs] = scores.squeeze(1).half()
            j += bs
        assert j == score_matrix.size(1)
    if was_training:
        model.train()
    return score_matrix


@torch.no_grad()
def evaluate(model, captions, arena, txt2img, img2txts, txt_ids=None):
    """itm.py:492-512: score, all-gather the row shards, recall on rank 0 (other ranks return {})."""
    world, rank = D.size(), D.rank()
    local = inference(model, captions, arena, rank, world)
    all_score = D.allgather_rows(local)
    txt_ids = list(range(len(captions))) if txt_ids is None else list(txt_ids)
    all_txt_ids = [t for r in range(world) for t in txt_ids[r::world]]      # the order the shards concatenate in
    assert tuple(all_score.shape) == (len(all_txt_ids), arena.num_images)
    if rank != 0:
        return {}
    return itm_eval(all_score, all_txt_ids, arena.img_ids, txt2img, img2txts)


@torch.no_grad()
def itm_eval(score_matrix, txt_ids, img_ids, txt2img, img2txts, column_only=False):
    """eval/itm.py:6-53 with the per-image Python loop turned into device ops (same numbers).

    Reference quirk kept by default: eval/itm.py:16-20 thresholds the WHOLE output of `.nonzero()` -- [row, rank]
    pairs -- so a hit in caption row i also counts once more for every recall level above i (upstream UNITER
    slices `[:, 1]`; this reference does not).  `column_only=True` gives the intended image-retrieval recall."""
    dev = score_matrix.device
    k = min(10, score_matrix.size(1))
    img2j = {i: j for j, i in enumerate(img_ids)}
    # image retrieval: rank of the ground-truth image among the top-10 of each caption row
    _, rank_txt = score_matrix.topk(k, dim=1)
    gt_img_j = torch.tensor([img2j[txt2img[t]] for t in txt_ids], dtype=torch.long, device=dev).unsqueeze(1)
    hit = rank_txt == gt_img_j
    pos = torch.where(hit.any(1), hit.float().argmax(1), torch.full((len(txt_ids),), 10, device=dev))
    hit_rows = hit.any(1).nonzero().squeeze(1)
    ir = [float(((pos < r).sum() + (0 if column_only else (hit_rows < r).sum())).item()) / len(txt_ids)
          for r in (1, 5, 10)]
    # text retrieval: best rank of any ground-truth caption among the top-10 of each image column
    k0 = min(10, score_matrix.size(0))
    _, rank_img = score_matrix.topk(k0, dim=0)                                   # [k0, n_img]
    txt2i = {t: i for i, t in enumerate(txt_ids)}
    is_gt = torch.zeros(score_matrix.shape, dtype=torch.bool, device=dev)
    rows = [txt2i[t] for img in img_ids for t in img2txts[img] if t in txt2i]
    cols = [j for j, img in enumerate(img_ids) for t in img2txts[img] if t in txt2i]
    if rows:
        is_gt[torch.tensor(rows, device=dev), torch.tensor(cols, device=dev)] = True
    hit_i = is_gt.gather(0, rank_img)                                            # [k0, n_img]
    pos_i = torch.where(hit_i.any(0), hit_i.float().argmax(0), torch.full((len(img_ids),), 10, device=dev))
    tr = [float((pos_i < r).sum().item()) / len(img_ids) for r in (1, 5, 10)]
    tr_mean, ir_mean = sum(tr) / 3, sum(ir) / 3
    return {"txt_r1": tr[0], "txt_r5": tr[1], "txt_r10": tr[2], "txt_r_mean": tr_mean,
            "img_r1": ir[0], "img_r5": ir[1], "img_r10": ir[2], "img_r_mean": ir_mean,
            "r_mean": (tr_mean + ir_mean) / 2}


@torch.no_grad()
def validate(model, val_loader):
    """itm.py:447-488: every batch is one caption against `mini_batch_size` images with the ground truth FIRST; the
    rank of index 0 inside the top-10 gives recall@1/5/10.  The reference pulls `rank.item()` to the host per batch;
    here the three counters live on the device and are read once."""
    import time
    was_training = model.training
    model.eval()
    st = time.time()
    counts, n_ex = None, 0
    for batch in val_loader:
        scores = model(batch, compute_loss=False).squeeze(1)
        _, indices = scores.topk(min(10, scores.numel()), dim=0)
        hit = indices == 0
        pos = torch.where(hit.any(), hit.float().argmax(), torch.full((), 10, device=scores.device))
        c = torch.stack([pos < 1, pos < 5, pos < 10]).long()
        counts = c if counts is None else counts + c
        n_ex += 1
    r = [0, 0, 0] if counts is None else [int(x) for x in counts.tolist()]
    n_all = sum(D.all_gather_list(n_ex))
    r = [sum(D.all_gather_list(x)) / max(n_all, 1) for x in r]
    tot = time.time() - st
    if was_training:
        model.train()
    return {"valid/ex_per_s": n_all / tot, "valid/recall_1": r[0], "valid/recall_5": r[1], "valid/recall_10": r[2]}


def hardest_per_group(scores, group, k):
    """For every distinct value of `group` (int64 [N]) the indices of its (up to) k largest `scores` (descending,
    ties by first occurrence).  Returns (sel [M] indices into scores, grouped by ascending group value, sizes dict
    is implied by group[sel]).  Works on any device: two stable sorts and a segmented position."""
    order = torch.argsort(scores, descending=True, stable=True)
    order = order[torch.argsort(group[order], stable=True)]          # by group, score-descending inside each group
    g = group[order]
    start = torch.ones_like(g, dtype=torch.bool)
    start[1:] = g[1:] != g[:-1]
    idx = torch.arange(g.numel(), device=g.device)
    first = torch.cummax(torch.where(start, idx, torch.zeros_like(idx)), 0)[0]
    return order[(idx - first) < k]


@torch.no_grad()
def get_hard_negs(model, loader, hard_negative_num=20, all_img_ids=None):
    """itm.py:385-445.  Each batch: one caption (`gt_txt_id`) scored against `neg_img_ids` images.
    txt2hardimgs[txt] = the `hard_negative_num` best-scoring images of that caption; img2hardtxts[img] = the
    `hard_negative_num` best-scoring captions of that image over ALL ranks (rank 0 only, like the reference; fewer
    when an image met fewer captions).  The reference calls topk(sorted=False), whose order is unspecified: lists
    here are in descending score order.  Scores stay on the device until the end (the reference calls .item() per
    pair); the per-image selection is two device sorts instead of a Python loop over images."""
    was_training = model.training
    model.eval()
    txt2hardimgs = {}
    img_index, txt_list, hard_idx = {}, [], []
    all_scores, all_img, all_txt = [], [], []
    for batch in loader:
        scores = model(batch, compute_loss=False).squeeze(-1).float()
        txt, imgs = batch["gt_txt_id"], batch["neg_img_ids"]
        assert scores.numel() == len(imgs)
        k = min(hard_negative_num, len(imgs))
        hard_idx.append(scores.topk(k)[1])                            # device indices, resolved after the loop
        t = len(txt_list)
        txt_list.append((txt, imgs))
        all_scores.append(scores)
        all_img.append(torch.tensor([img_index.setdefault(i, len(img_index)) for i in imgs], dtype=torch.long))
        all_txt.append(torch.full((len(imgs),), t, dtype=torch.long))
    for (txt, imgs), hi in zip(txt_list, hard_idx):
        txt2hardimgs[txt] = [imgs[i] for i in hi.tolist()]
    # hard texts per image: pairs from every rank
    local_imgs = sorted(img_index, key=img_index.get)
    if all_scores:
        sc = torch.cat(all_scores).cpu()
        pairs = (sc, torch.cat(all_img), torch.cat(all_txt))
    else:
        pairs = (torch.zeros(0), torch.zeros(0, dtype=torch.long), torch.zeros(0, dtype=torch.long))
    gathered = D.all_gather_list((pairs, local_imgs, [t for t, _ in txt_list]))
    if was_training:
        model.train()
    if D.rank() != 0:
        return txt2hardimgs, {}
    gimg, scs, gi, gt, txt_names = {}, [], [], [], []
    for (sc, im, tx), names, txts in gathered:
        remap = torch.tensor([gimg.setdefault(n, len(gimg)) for n in names], dtype=torch.long)
        scs.append(sc)
        gi.append(remap[im] if im.numel() else im)
        gt.append(tx + len(txt_names))
        txt_names.extend(txts)
    sc, gi, gt = torch.cat(scs), torch.cat(gi), torch.cat(gt)
    dev = all_scores[0].device if all_scores else "cpu"
    sel = hardest_per_group(sc.to(dev), gi.to(dev), hard_negative_num).cpu()
    names = sorted(gimg, key=gimg.get)
    img2hardtxts = {}
    for i, t in zip(gi[sel].tolist(), gt[sel].tolist()):
        img2hardtxts.setdefault(names[i], []).append(txt_names[t])
    for img in (all_img_ids if all_img_ids is not None else names):
        img2hardtxts.setdefault(img, [])
    return txt2hardimgs, img2hardtxts
