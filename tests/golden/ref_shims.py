"""Import shims that let the UNMODIFIED reference modules under /root/reference run on
CPU in the authoring container (SURVEY.md 8c lists and justifies each one).  Used only
by tests/golden/make_golden.py and oracle/validate_against_reference.py -- never at GPU
test / bench time (the reference does not travel to the GPU box).
"""
import itertools
import sys
import types

import torch

REF = "/root/reference"


def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def install(valid_token_ids=None):
    if "model.model" in sys.modules:
        return
    # (1) apex FusedLayerNorm -> torch LayerNorm (BASELINE.json configs[0])
    _mod("apex")
    _mod("apex.normalization")
    _mod("apex.normalization.fused_layer_norm", FusedLayerNorm=torch.nn.LayerNorm)
    # (5) packages the collate modules import but the path does not need
    hvd = _mod("horovod.torch", size=lambda: 1, rank=lambda: 0, local_rank=lambda: 0)
    _mod("horovod", torch=hvd)
    _mod("lmdb")
    lz4 = _mod("lz4")
    lz4.frame = _mod("lz4.frame", compress=None, decompress=None)
    _mod("msgpack_numpy", patch=lambda: None)
    tz = _mod("toolz")
    tz.sandbox = _mod("toolz.sandbox", unzip=lambda s: zip(*s))
    def partition_all(n, seq):
        seq = list(seq)
        return [tuple(seq[i:i + n]) for i in range(0, len(seq), n)]
    _mod("cytoolz", concat=itertools.chain.from_iterable, partition_all=partition_all, curry=lambda f: f)
    _mod("tensorboardX", SummaryWriter=object)
    # (2) const_variable downloads xlm-roberta-base at import; only the unused vis_cls head needs it
    pkg = types.ModuleType("model")
    pkg.__path__ = [REF + "/model"]
    sys.modules["model"] = pkg
    _mod("model.const_variable", XLMR_TOKER=None, LABEL2TOKEN_MATRIX=None,
         VALID_XLMR_TOKEN_IDS=list(valid_token_ids or range(16)))
    # data/__init__.py pulls data/mlm.py (tokenizer download): register a bare package instead
    dpk = types.ModuleType("data")
    dpk.__path__ = [REF + "/data"]
    sys.modules["data"] = dpk
    opk = types.ModuleType("optim")
    opk.__path__ = [REF + "/optim"]
    sys.modules["optim"] = opk
    # data/loader.py imports utils.distributed (horovod helpers): register the package so it resolves from the
    # reference tree without executing anything else of utils/
    upk = types.ModuleType("utils")
    upk.__path__ = [REF + "/utils"]
    sys.modules["utils"] = upk
    import model.ot as ot
    # (4) ot.trace builds a uint8 eye for masked_select, which torch>=2 rejects; same maths:
    ot.trace = lambda x: x.diagonal(dim1=-2, dim2=-1).sum(-1)


def load():
    install()
    import model.model as mm
    import model.itm as mi
    import model.ot as mo
    import model.layer as ml
    import optim.adamw as oa
    import optim.misc as om
    import optim.sched as osch
    import data.data as dd
    import data.itm as di
    import data.mrm as dm
    import data.sampler as dsamp
    import data.loader as dload
    import importlib.util
    spec = importlib.util.spec_from_file_location("ref_eval_itm", REF + "/eval/itm.py")
    ev = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ev)
    return types.SimpleNamespace(model=mm, itm=mi, ot=mo, layer=ml, adamw=oa, misc=om, sched=osch,
                                 data=dd, data_itm=di, data_mrm=dm, eval_itm=ev, sampler=dsamp, loader=dload)
