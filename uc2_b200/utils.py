"""Small host helpers mirrored from the reference's utils/misc.py and utils/const.py."""
import random

import numpy as np
import torch

IMG_DIM = 2048
IMG_LABEL_DIM = 1601
BUCKET_SIZE = 8192


class NoOp(object):
    """useful for distributed training No-Ops (utils/misc.py:14-20)"""
    def __getattr__(self, name):
        return self.noop

    def noop(self, *args, **kwargs):
        return


def set_dropout(model, drop_p):
    """utils/misc.py:54-60"""
    for name, module in model.named_modules():
        if isinstance(module, torch.nn.Dropout):
            if module.p != drop_p:
                module.p = drop_p


def set_random_seed(seed):
    """utils/misc.py:63-67"""
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    torch.cuda.manual_seed_all(seed)
