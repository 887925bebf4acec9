"""uc2_b200: the UC2 (UNITER-derived) cross-modal encoder path on B200 (sm_100a).

Modules, named after what they replace in the reference tree:

    model, itm          model/model.py, model/itm.py      (same classes, forward signatures, state_dict names)
    ot                  model/ot.py
    optim               optim/adamw.py, optim/misc.py, optim/sched.py
    distributed         utils/distributed.py              (NCCL instead of horovod)
    batch, loader       data/*.py collates, data/sampler.py, data/loader.py
    sampling            token / region masking and negative sampling draws (data/mlm.py, data/mrm.py, data/itm.py)
    device_batch        the same collates built on the device from an HBM-resident feature arena
    datasets            in-memory task datasets over that arena (data/itm.py, data/mlm.py, data/mrm.py item logic)
    retrieval           itm.py evaluate / inference / validate / get_hard_negs, eval/itm.py
    train, pretrain_loop, validate      the loops of pretrain.py and itm.py
    save, utils         utils/save.py, utils/misc.py
    functional, arena, dropout, _lib    autograd nodes over the C ABI (include/uc2_b200.h -> libuc2_b200.so)

Nothing here runs without the CUDA library: there is no CPU or eager fallback.
"""
