"""Small host helpers with the names the reference's scripts import (utils/misc.py, utils/const.py)."""
import random

import numpy as np
import torch

IMG_DIM = 2048
IMG_LABEL_DIM = 1601
BUCKET_SIZE = 8192


class NoOp(object):
    """Stand-in for a progress bar / logger on non-zero ranks: every attribute is a callable that does nothing
    (the role of utils/misc.py:14-20)."""

    def __getattr__(self, name):
        return lambda *args, **kwargs: None


def set_dropout(model, drop_p):
    """Point every nn.Dropout of `model` at probability drop_p (utils/misc.py:54-60).  The fused kernels read the
    probabilities from these modules at forward time (model.py: _dropout_cfg)."""
    p = float(drop_p)
    for module in model.modules():
        if isinstance(module, torch.nn.Dropout):
            module.p = p


def set_random_seed(seed):
    """Seed python, numpy and torch (CPU and every CUDA device): utils/misc.py:63-67."""
    for seeder in (random.seed, np.random.seed, torch.manual_seed, torch.cuda.manual_seed_all):
        seeder(seed)
