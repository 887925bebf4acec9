"""In-memory datasets over an HBM-resident feature arena: what sits between the text / image stores and the collates.

The reference's datasets (data/data.py DetectFeatTxtTokDataset and its task subclasses in data/itm.py, data/mlm.py,
data/mrm.py) read LMDB and return per-sample TENSORS of region features, which the collate pads on the host.  The
LMDB storage layer is out of scope; here a `TextDB` holds the token ids and the caption -> image map in memory, the
features live in a `device_batch.FeatureArena`, and a dataset item is a few integers (token ids, the image's arena
index, masks).  The matching `collate` methods hand a list of items to `DeviceCollator`, so the padded batch is built
on the device.  Sampling order and random draws follow the reference (uc2_b200/sampling.py), sample lengths
`lens` feed `loader.TokenBucketSampler` exactly like `dset.lens` does there.
"""
import numpy as np
import torch

from . import sampling as S


class TextDB(object):
    """Stand-in for TxtTokLmdb (data/data.py:193-231).  `examples`: {id: {'input_ids': [int], 'img_fname': str}}.
    Captions longer than max_txt_len are dropped, as the LMDB wrapper does with its id2len filter."""

    def __init__(self, examples, max_txt_len=60, cls_=0, sep=2, mask=250001, v_range=(5, 250001)):
        self.cls_, self.sep, self.mask, self.v_range = cls_, sep, mask, tuple(v_range)
        self.id2len = {i: len(ex["input_ids"]) for i, ex in examples.items()
                       if max_txt_len == -1 or len(ex["input_ids"]) <= max_txt_len}
        self.examples = examples
        self.ids = list(self.id2len)

    def __getitem__(self, id_):
        return self.examples[id_]

    def combine_inputs(self, *inputs):
        out = [self.cls_]
        for ids in inputs:
            out.extend(list(ids) + [self.sep])
        return torch.tensor(out)

    @property
    def txt2img(self):
        return {i: self.examples[i]["img_fname"] for i in self.ids}

    @property
    def img2txts(self):
        out = {}
        for i in self.ids:
            out.setdefault(self.examples[i]["img_fname"], []).append(i)
        return out


class ImageIndex(object):
    """name -> arena index / number of boxes (the role of DetectFeatLmdb.name2nbb)."""

    def __init__(self, arena, names):
        assert len(names) == len(arena)
        self.arena = arena
        self.index = {n: i for i, n in enumerate(names)}
        self.name2nbb = {n: arena.nbb[i] for i, n in enumerate(names)}


class _Base(object):
    """DetectFeatTxtTokDataset data/data.py:291-318: ids (sharded over ranks like hvd.rank()::hvd.size()) and
    lens = text length + number of boxes."""

    def __init__(self, txt_db, img_index, rank=0, world=1):
        self.txt_db, self.img = txt_db, img_index
        ids = txt_db.ids[rank::world]
        self.ids = ids
        self.txt_lens = [txt_db.id2len[i] for i in ids]
        self.lens = [tl + img_index.name2nbb[txt_db[i]["img_fname"]] for tl, i in zip(self.txt_lens, ids)]

    def __len__(self):
        return len(self.ids)

    def _example(self, i):
        return self.txt_db[self.ids[i]]

    def _img(self, name):
        return self.img.index[name]


class MlmDataset(_Base):
    """data/mlm.py:452-500: item = (masked ids incl. specials, labels, image)."""

    def __getitem__(self, i):
        ex = self._example(i)
        ids, labels = S.create_mlm_io(ex["input_ids"], self.txt_db.v_range, self.txt_db.mask, self.txt_db.cls_,
                                      self.txt_db.sep)
        return ids, labels, self._img(ex["img_fname"])

    @staticmethod
    def collate(dc, items):
        return dc.mlm([(a, b) for a, b, _ in items], [k for _, _, k in items])


class MrfrDataset(_Base):
    """data/mrm.py:42-70: item = (ids, image, region mask)."""

    def __init__(self, mask_prob, *args, **kw):
        super().__init__(*args, **kw)
        self.mask_prob = mask_prob

    def __getitem__(self, i):
        ex = self._example(i)
        name = ex["img_fname"]
        return (self.txt_db.combine_inputs(ex["input_ids"]), self._img(name),
                S.get_img_mask(self.mask_prob, self.img.name2nbb[name]))

    @staticmethod
    def collate(dc, items):
        return dc.mrfr([a for a, _, _ in items], [k for _, k, _ in items], [m for _, _, m in items])


class MrcDataset(MrfrDataset):
    """data/mrm.py:221-255: as MRFR; the soft labels are read from the arena by the collator."""

    @staticmethod
    def collate(dc, items):
        return dc.mrc([a for a, _, _ in items], [k for _, k, _ in items], [m for _, _, m in items])


class ItmDataset(_Base):
    """data/itm.py:150-199: every epoch each caption keeps its image (label 1) or gets a random other one
    (label 0, probability neg_sample_p); item = (ids, image, label)."""

    def __init__(self, txt_db, img_index, neg_sample_p=0.5, rank=0, world=1):
        super().__init__(txt_db, img_index, rank, world)
        self.all_imgs = list(set(txt_db[i]["img_fname"] for i in self.ids))
        self.neg_sample_p = neg_sample_p
        self.new_epoch()

    def new_epoch(self):
        self.labels = np.random.choice([0, 1], size=len(self.ids), p=[self.neg_sample_p, 1 - self.neg_sample_p])
        self.lens, self.train_imgs = [], []
        for i, tl in enumerate(self.txt_lens):
            name = self._example(i)["img_fname"]
            if self.labels[i] == 0:
                name = S.sample_negative(self.all_imgs, [name], 1)[0]
            self.train_imgs.append(name)
            self.lens.append(tl + self.img.name2nbb[name])

    def __getitem__(self, i):
        return (self.txt_db.combine_inputs(self._example(i)["input_ids"]), self._img(self.train_imgs[i]),
                int(self.labels[i]))

    @staticmethod
    def collate(dc, items, with_ot=True):
        return dc.itm([a for a, _, _ in items], [k for _, k, _ in items], [t for _, _, t in items], with_ot=with_ot)


class ItmRankDataset(_Base):
    """data/itm.py:355-395: item = 1 positive + n wrong-image + n wrong-caption pairs, each (ids, image)."""

    def __init__(self, txt_db, img_index, neg_sample_size=1, rank=0, world=1):
        assert neg_sample_size > 0, "ItmRankDataset need at least 1 negative sample"
        super().__init__(txt_db, img_index, rank, world)
        t2i = txt_db.txt2img
        self.txt2img = {i: t2i[i] for i in self.ids}
        self.img2txts = {}
        for i, img in self.txt2img.items():
            self.img2txts.setdefault(img, []).append(i)
        self.img_name_list = list(self.img2txts)
        self.neg_sample_size = neg_sample_size

    def __getitem__(self, i):
        gt_txt = self.ids[i]
        gt_img = self.txt2img[gt_txt]
        pairs = S.rank_id_pairs(gt_txt, gt_img, self.img_name_list, self.ids, self.img2txts[gt_img],
                                self.neg_sample_size)
        assert len(pairs) == 1 + 2 * self.neg_sample_size
        return [(self.txt_db.combine_inputs(self.txt_db[t]["input_ids"]), self._img(im)) for t, im in pairs]

    def collate(self, dc, items):
        """xlmr_itm_rank_collate data/itm.py:615-643 takes ONE item (batch_size 1 of 1 + 2n pairs); several items are
        concatenated group after group, which is what the triplet loss's view(-1, sample_size) expects."""
        flat = [p for it in items for p in it]
        return dc.itm_rank([a for a, _ in flat], [k for _, k in flat], 1 + 2 * self.neg_sample_size)


class BatchLoader(object):
    """Iterable of collated batches: what `DataLoader(dset, batch_sampler=sampler, collate_fn=...)` of
    build_dataloader (pretrain.py:82-93) yields, in the calling process -- items are a few integers, there is nothing
    for worker processes to do.  `collate(items) -> batch` is typically `lambda items: Dataset.collate(dc, items)`."""

    def __init__(self, dataset, sampler, collate):
        self.dataset, self.sampler, self.collate = dataset, sampler, collate

    def __iter__(self):
        for idx in self.sampler:
            yield self.collate([self.dataset[i] for i in idx])
