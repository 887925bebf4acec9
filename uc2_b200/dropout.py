"""Counter-based dropout masks shared by every fused kernel on the path.

The reference applies nn.Dropout at five places (model/model.py:334 text embeddings, 363 image embeddings;
model/layer.py:94 attention probabilities, 113 attention output, 154 FFN output).  A fused kernel cannot keep a
mask tensor around between forward and backward without paying HBM for it, so both passes REGENERATE the mask from
(site key, element index):

    keep(idx)  <=>  half(lowbias32((idx >> 1) ^ key), idx & 1) >= thresh,   thresh = round(p * 65536),  scale = 1 / (1 - p)

(one 32-bit mix serves the two neighbouring elements 2j and 2j + 1: its low 16 bits decide the even one, its high 16
bits the odd one -- every kernel holds such pairs along its fastest index and pays one mix per pair)

`lowbias32` is Chris Wellons' 32-bit integer mixer (public domain); csrc/common.cuh has the same function.  The
site key mixes the user seed, a per-forward counter, the layer and the site, so every step / layer / site draws an
independent mask; the element index is row * width + column (attention: q * S + k under a per-(batch, head) key).
This file is the host-side mirror: the modules derive keys with it, and the tests rebuild the exact masks with
numpy to drive the oracle.
"""
import numpy as np

M32 = 0xFFFFFFFF
SITE_ATTN, SITE_OUT1, SITE_OUT2, SITE_EMB = 0, 1, 2, 3


def lowbias32(x):
    x &= M32
    x ^= x >> 16
    x = (x * 0x7FEB352D) & M32
    x ^= x >> 15
    x = (x * 0x846CA68B) & M32
    x ^= x >> 16
    return x


def site_key(seed, counter, layer, site):
    """32-bit key of one dropout site of one forward pass."""
    k = lowbias32((seed & M32) ^ 0x9E3779B9)
    k = lowbias32(k ^ ((seed >> 32) & M32))
    k = lowbias32(k ^ ((counter * 0x85EBCA6B) & M32))
    return lowbias32(k ^ (((layer * 4 + site + 1) * 0xC2B2AE35) & M32))


def head_key(key, bh):
    """Attention: key of (batch * 12 + head) under a layer's attention-site key."""
    return lowbias32(key ^ ((bh * 0x9E3779B9 + 0x7F4A7C15) & M32))


def thresh_of(p):
    if not 0.0 <= p < 1.0:
        raise ValueError("dropout probability has to be in [0, 1), got {}".format(p))
    return int(round(p * 65536.0))


def scale_of(p):
    t = thresh_of(p)
    return 65536.0 / (65536.0 - t) if t else 1.0      # 1 / (1 - effective p), effective p = thresh / 65536


def lowbias32_np(x):
    x = x.astype(np.uint32)
    x ^= x >> np.uint32(16)
    x = (x.astype(np.uint64) * np.uint64(0x7FEB352D)).astype(np.uint32)
    x ^= x >> np.uint32(15)
    x = (x.astype(np.uint64) * np.uint64(0x846CA68B)).astype(np.uint32)
    x ^= x >> np.uint32(16)
    return x


ATTN_CA, ATTN_CB = 0x9E3779B1, 0x85EBCA77


def attn_keep_mask_np(hkey, S, thresh):
    """[S, S] boolean keep mask (query, key) of one (batch, head) for the tcgen05 attention kernels
    (csrc/common.cuh, "Attention-probability dropout"): one lowbias32 per 16 x 16 block, element (i, j) takes
    e = h_block * CA^(i & 15) * CB^(j & 15) mod 2^32 and is kept iff e >= thresh << 16."""
    i = np.arange(S, dtype=np.uint32)
    blk = ((i[:, None] >> np.uint32(4)) << np.uint32(16)) | (i[None, :] >> np.uint32(4))
    h = lowbias32_np(blk ^ np.uint32(hkey)).astype(np.uint64)
    pa = np.array([pow(ATTN_CA, k, 1 << 32) for k in range(16)], dtype=np.uint64)
    pb = np.array([pow(ATTN_CB, k, 1 << 32) for k in range(16)], dtype=np.uint64)
    m = np.uint64(M32)
    e = (((h * pa[i & np.uint32(15)][:, None]) & m) * pb[i & np.uint32(15)][None, :]) & m
    return e >= np.uint64(int(thresh) << 16)


def keep_mask_np(key, n, thresh, idx=None):
    """Boolean keep mask of elements 0..n-1 (or of the given uint32 index array) under `key`."""
    if idx is None:
        idx = np.arange(n, dtype=np.uint64)
    idx = (np.asarray(idx).astype(np.uint64) & np.uint64(M32)).astype(np.uint32)
    h = lowbias32_np((idx >> np.uint32(1)) ^ np.uint32(key))
    half = np.where((idx & np.uint32(1)) == 1, h >> np.uint32(16), h & np.uint32(0xFFFF))
    return half >= np.uint32(thresh)
