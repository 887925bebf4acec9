"""B200-native drop-in modules for the UC2 cross-modal encoder path.

Module tree, parameter names, constructor and forward signatures mirror the reference so that
pretrain.py / itm.py and reference checkpoints work unchanged:

  VLXLMRModel / UniterModel                 model/model.py:385-458 / 1067-1140
  VLXLMRForPretraining / UniterForPretraining   model/model.py:460-775 / 1172-1501

The torch modules below only HOLD parameters (nn.Linear / nn.Embedding / nn.LayerNorm objects are never
called); all arithmetic runs in the sm_100a kernels of libuc2_b200.so through uc2_b200.functional.
There is no CPU path: inputs must live on a CUDA device.
"""
import copy
import logging
import weakref
from collections import defaultdict

import torch
from torch import nn

from . import _lib
from . import functional as Fn
from .arena import ParamArena
from .config import UC2Config

logger = logging.getLogger(__name__)
LayerNorm = nn.LayerNorm


# --------------------------------------------------------------------------------------------------
# family switches
# --------------------------------------------------------------------------------------------------
class _Family(object):
    def __init__(self, name):
        self.name = name
        if name == "vlxlmr":
            self.prefix = "roberta."
            self.type_emb = "embeddings.new_token_type_embeddings.weight"
            self.derive_positions = True
        else:
            self.prefix = "bert."
            self.type_emb = "embeddings.token_type_embeddings.weight"
            self.derive_positions = False

    def bind(self, cfg):
        if self.name == "vlxlmr":
            self.word_pad = self.pos_pad = int(cfg.pad_token_id)
        else:
            self.word_pad, self.pos_pad = 0, -1
        return self

    def emb_eps(self, cfg):
        return float(cfg.layer_norm_eps) if self.name == "vlxlmr" else 1e-12


# --------------------------------------------------------------------------------------------------
# parameter holders with the reference's attribute names
# --------------------------------------------------------------------------------------------------
class _TextEmbeddings(nn.Module):
    def __init__(self, config, family):
        super().__init__()
        H = config.hidden_size
        if family == "vlxlmr":
            self.padding_idx = config.pad_token_id
            self.word_embeddings = nn.Embedding(config.vocab_size, H, padding_idx=config.pad_token_id)
            self.position_embeddings = nn.Embedding(config.max_position_embeddings, H, padding_idx=config.pad_token_id)
            self.new_token_type_embeddings = nn.Embedding(config.type_vocab_size, H)
            self.LayerNorm = LayerNorm(H, eps=config.layer_norm_eps)
        else:
            self.word_embeddings = nn.Embedding(config.vocab_size, H, padding_idx=0)
            self.position_embeddings = nn.Embedding(config.max_position_embeddings, H)
            self.token_type_embeddings = nn.Embedding(config.type_vocab_size, H)
            self.LayerNorm = LayerNorm(H, eps=1e-12)
        self.dropout = nn.Dropout(config.hidden_dropout_prob)


class _ImageEmbeddings(nn.Module):
    def __init__(self, config, img_dim, family):
        super().__init__()
        H = config.hidden_size
        eps = config.layer_norm_eps if family == "vlxlmr" else 1e-12
        self.img_linear = nn.Linear(img_dim, H)
        self.img_layer_norm = LayerNorm(H, eps=eps)
        self.pos_layer_norm = LayerNorm(H, eps=eps)
        self.pos_linear = nn.Linear(7, H)
        self.mask_embedding = nn.Embedding(2, img_dim, padding_idx=0)
        self.LayerNorm = LayerNorm(H, eps=eps)
        self.dropout = nn.Dropout(config.hidden_dropout_prob)


class _SelfAttention(nn.Module):
    def __init__(self, config):
        super().__init__()
        H = config.hidden_size
        self.query, self.key, self.value = nn.Linear(H, H), nn.Linear(H, H), nn.Linear(H, H)
        self.dropout = nn.Dropout(config.attention_probs_dropout_prob)


class _SelfOutput(nn.Module):
    def __init__(self, config, in_dim):
        super().__init__()
        self.dense = nn.Linear(in_dim, config.hidden_size)
        self.LayerNorm = LayerNorm(config.hidden_size, eps=1e-12)
        self.dropout = nn.Dropout(config.hidden_dropout_prob)


class _Attention(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.self = _SelfAttention(config)
        self.output = _SelfOutput(config, config.hidden_size)


class _Intermediate(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.dense = nn.Linear(config.hidden_size, config.intermediate_size)


class BertLayer(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.attention = _Attention(config)
        self.intermediate = _Intermediate(config)
        self.output = _SelfOutput(config, config.intermediate_size)


class _Encoder(nn.Module):
    def __init__(self, config):
        super().__init__()
        layer = BertLayer(config)
        self.layer = nn.ModuleList([copy.deepcopy(layer) for _ in range(config.num_hidden_layers)])


class BertPooler(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.dense = nn.Linear(config.hidden_size, config.hidden_size)
        self.activation = nn.Tanh()

    def forward(self, hidden_states):
        root = _root_of(self)
        arena = root._arena()
        pre = _name_of(root, self)
        return Fn.PoolerFn.apply(hidden_states, arena, pre + "dense.weight", pre + "dense.bias")


# --------------------------------------------------------------------------------------------------
# base class: config, init, from_pretrained, arena ownership
# --------------------------------------------------------------------------------------------------
def _root_of(module):
    r = getattr(module, "_uc2_root", None)
    root = r() if r is not None else None
    return root if root is not None else module


def _name_of(root, module):
    for n, m in root.named_modules():
        if m is module:
            return n + "." if n else ""
    raise RuntimeError("module is not part of its root")


def _adopt(root):
    for m in root.modules():
        if m is not root:
            object.__setattr__(m, "_uc2_root", weakref.ref(root))


class UC2PreTrainedModel(nn.Module):
    """Mirror of VLXLMRPreTrainedModel / UniterPreTrainedModel (model/model.py:145-278, 873-968)."""
    family_name = "vlxlmr"

    def __init__(self, config, *inputs, **kwargs):
        super().__init__()
        if not isinstance(config, UC2Config):
            raise ValueError(
                "Parameter config in `{}(config)` should be an instance of class `UC2Config`.".format(
                    self.__class__.__name__))
        config.check_kernel_support()
        self.config = config
        object.__setattr__(self, "_uc2_arena_obj", None)

    def init_weights(self, module):
        if isinstance(module, (nn.Linear, nn.Embedding)):
            module.weight.data.normal_(mean=0.0, std=self.config.initializer_range)
        elif isinstance(module, LayerNorm):
            module.bias.data.zero_()
            module.weight.data.fill_(1.0)
        if isinstance(module, nn.Linear) and module.bias is not None:
            module.bias.data.zero_()

    def state_dict(self, *args, **kwargs):
        """The reference's nn.Module.state_dict(); with a deferred optimizer (optim.AdamW(lazy_rows=True)) the postponed
        word-embedding row updates are applied first, so the tensors handed out are what eager AdamW would hold."""
        a = object.__getattribute__(_root_of(self), "_uc2_arena_obj") if hasattr(_root_of(self), "_uc2_arena_obj") else None
        lazy = getattr(a, "lazy", None) if a is not None else None
        if lazy is not None:
            lazy.catch_up_all()
        return super().state_dict(*args, **kwargs)

    # ---- arena ----------------------------------------------------------------------------------
    def _arena(self):
        root = _root_of(self)
        if root is not self:
            return root._arena()
        a = self._uc2_arena_obj
        p0 = next(self.parameters())
        if not p0.is_cuda:
            raise RuntimeError("uc2_b200 has no CPU path: move the model to a CUDA device (model.cuda()) first")
        if a is None or not a.intact():
            a = ParamArena(self, p0.device)
            object.__setattr__(self, "_uc2_arena_obj", a)
        a.sync_shadow()
        return a

    def _dropout_cfg(self):
        """(p_hidden, p_attn, seed, counter) for this forward pass, or None.  The probabilities are read from the
        nn.Dropout modules the reference declares (so utils.set_dropout / model.eval() act as usual); every
        hidden-state dropout has to share one p and every attention dropout one p, as in every UC2 config."""
        if not self.training:
            return None
        hid, att = set(), set()
        for n, m in self.named_modules():
            if isinstance(m, nn.Dropout):
                (att if n.endswith("attention.self.dropout") else hid).add(float(m.p))
        if len(hid) > 1 or len(att) > 1:
            raise NotImplementedError("per-module dropout probabilities are not supported: hidden {} attention {}"
                                      .format(sorted(hid), sorted(att)))
        p_h, p_a = (hid.pop() if hid else 0.0), (att.pop() if att else 0.0)
        if p_h == 0.0 and p_a == 0.0:
            return None
        c = getattr(self, "_uc2_drop_counter", 0) + 1
        object.__setattr__(self, "_uc2_drop_counter", c)
        return (p_h, p_a, int(torch.initial_seed()), c)

    @classmethod
    def from_pretrained(cls, config_file, state_dict, load_embedding_only=False, load_layer=None, *inputs, **kwargs):
        """Same contract as model/model.py:174-278: build from a config json, then load a state dict
        (gamma/beta renames, optional 'roberta.bert.' prefix, partial XLM-R loading)."""
        config = UC2Config.from_json_file(config_file)
        logger.info("Model config {}".format(config))
        model = cls(config, *inputs, **kwargs)
        state_dict = dict(state_dict)
        if load_embedding_only or load_layer:
            for key in list(state_dict.keys()):
                drop = ("roberta.embeddings" not in key) if load_embedding_only else (
                    "roberta.encoder" in key and int(key.split(".")[3]) > load_layer)
                if drop:
                    state_dict["not_load." + key] = state_dict.pop(key)
        else:
            for key in list(state_dict.keys()):
                new_key = key.replace("gamma", "weight") if "gamma" in key else key
                new_key = new_key.replace("beta", "bias") if "beta" in new_key else new_key
                if new_key != key:
                    state_dict[new_key] = state_dict.pop(key)
        if any(s.startswith("roberta.bert.") for s in state_dict.keys()):
            state_dict = {(k[len("roberta.bert."):] if k.startswith("roberta.bert.") else k): v
                          for k, v in state_dict.items()}
        missing, unexpected = model.load_state_dict(state_dict, strict=False)
        if missing:
            logger.info("Weights of {} not initialized from pretrained model: {}".format(cls.__name__, missing))
        if unexpected:
            logger.info("Weights from pretrained model not used in {}: {}".format(cls.__name__, unexpected))
        return model


# --------------------------------------------------------------------------------------------------
# encoder core
# --------------------------------------------------------------------------------------------------
class _EncoderModel(UC2PreTrainedModel):
    """Joint vision-language encoder: embeddings -> pack -> BertLayer stack (+ pooler)."""

    def __init__(self, config, img_dim):
        super().__init__(config)
        fam = self.family_name
        self.embeddings = _TextEmbeddings(config, fam)
        self.img_embeddings = _ImageEmbeddings(config, img_dim, fam)
        self.encoder = _Encoder(config)
        self.pooler = BertPooler(config)
        self.apply(self.init_weights)
        self.family = _Family(fam).bind(config)
        object.__setattr__(self, "_anchor", None)
        _adopt(self)

    # names inside the root module ------------------------------------------------------------------
    @property
    def prefix(self):
        root = _root_of(self)
        return _name_of(root, self)

    def layer_param_names(self):
        pre = self.prefix
        out = []
        for l in range(self.config.num_hidden_layers):
            q = pre + f"encoder.layer.{l}."
            for n in ("query", "key", "value"):
                out += [q + f"attention.self.{n}.weight", q + f"attention.self.{n}.bias"]
            out += [q + "attention.output.dense.weight", q + "attention.output.dense.bias",
                    q + "attention.output.LayerNorm.weight", q + "attention.output.LayerNorm.bias",
                    q + "intermediate.dense.weight", q + "intermediate.dense.bias",
                    q + "output.dense.weight", q + "output.dense.bias",
                    q + "output.LayerNorm.weight", q + "output.LayerNorm.bias"]
        return out

    def _layer_structs(self, arena, cls, wptr, fptr):
        L = self.config.num_hidden_layers
        arr = (cls * L)()
        pre = self.prefix
        for l in range(L):
            q = pre + f"encoder.layer.{l}."
            s = arr[l]
            s.w_qkv = wptr(q + "attention.self.query.weight")
            s.b_qkv = fptr(q + "attention.self.query.bias")
            s.w_o, s.b_o = wptr(q + "attention.output.dense.weight"), fptr(q + "attention.output.dense.bias")
            s.ln1_w, s.ln1_b = fptr(q + "attention.output.LayerNorm.weight"), fptr(q + "attention.output.LayerNorm.bias")
            s.w_ffn1, s.b_ffn1 = wptr(q + "intermediate.dense.weight"), fptr(q + "intermediate.dense.bias")
            s.w_ffn2, s.b_ffn2 = wptr(q + "output.dense.weight"), fptr(q + "output.dense.bias")
            s.ln2_w, s.ln2_b = fptr(q + "output.LayerNorm.weight"), fptr(q + "output.LayerNorm.bias")
        return arr

    def layer_weight_structs(self, arena):
        return self._layer_structs(arena, _lib.LayerWeights, arena.sp, arena.mp)

    def layer_grad_structs(self, arena):
        return self._layer_structs(arena, _lib.LayerGrads, arena.gp, arena.gp)

    # reference-named helpers -----------------------------------------------------------------------
    def _compute_img_txt_embeddings(self, input_ids, position_ids, img_feat, img_pos_feat, gather_index,
                                    img_masks=None, txt_type_ids=None, img_type_ids=None):
        """Packed embedding output only (no grad); model/model.py:412-425."""
        am = torch.ones(gather_index.shape, dtype=torch.long, device=gather_index.device)
        with torch.no_grad():
            x0, _, _ = Fn.encoder_forward(self, self._arena(), input_ids, self._pos(position_ids), img_feat,
                                          img_pos_feat, am, gather_index, img_masks, save=False, embed_only=True,
                                          dropout=_root_of(self)._dropout_cfg() if self.training else None)
        return x0

    def _pos(self, position_ids):
        return position_ids

    def forward(self, input_ids, position_ids, img_feat, img_pos_feat, attention_mask, gather_index=None,
                img_masks=None, output_all_encoded_layers=True, txt_type_ids=None, img_type_ids=None):
        if txt_type_ids is not None or img_type_ids is not None:
            raise NotImplementedError("explicit token type ids are not used by any UC2 call site and are not "
                                      "implemented (text = type 0, regions = type 1)")
        root = _root_of(self)
        arena = self._arena()
        kw = dict(input_ids=input_ids, position_ids=position_ids, img_feat=img_feat, img_pos_feat=img_pos_feat,
                  attention_mask=attention_mask, gather_index=gather_index, img_masks=img_masks,
                  dropout=root._dropout_cfg() if self.training else None)
        if torch.is_grad_enabled():
            if output_all_encoded_layers:
                raise NotImplementedError("output_all_encoded_layers=True is only available under torch.no_grad() "
                                          "(every training call site of the reference passes False)")
            if self._anchor is None or self._anchor.device != attention_mask.device:
                object.__setattr__(self, "_anchor", torch.zeros(1, device=attention_mask.device, requires_grad=True))
            return Fn.EncoderFn.apply(self._anchor, self, arena, kw)
        _, outs, _ = Fn.encoder_forward(self, arena, save=False, keep_all=output_all_encoded_layers, **kw)
        return outs if output_all_encoded_layers else outs[-1]


class VLXLMRModel(_EncoderModel):
    family_name = "vlxlmr"


class UniterModel(_EncoderModel):
    family_name = "uniter"


# --------------------------------------------------------------------------------------------------
# head parameter holders
# --------------------------------------------------------------------------------------------------
class RobertaLMHead(nn.Module):
    """model/layer.py:236-265; decoder weight tied to the word embeddings, decoder.bias tied to bias."""

    def __init__(self, config, emb_weight):
        super().__init__()
        self.dense = nn.Linear(config.hidden_size, config.hidden_size)
        self.layer_norm = LayerNorm(config.hidden_size, eps=config.layer_norm_eps)
        self.decoder = nn.Linear(emb_weight.size(1), emb_weight.size(0), bias=False)
        self.decoder.weight = emb_weight
        self.bias = nn.Parameter(torch.zeros(emb_weight.size(0)))
        self.decoder.bias = self.bias


class _BertTransform(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.dense = nn.Linear(config.hidden_size, config.hidden_size)
        self.LayerNorm = LayerNorm(config.hidden_size, eps=1e-12)


class _BertLMPredictionHead(nn.Module):
    def __init__(self, config, emb_weight):
        super().__init__()
        self.transform = _BertTransform(config)
        self.decoder = nn.Linear(emb_weight.size(1), emb_weight.size(0), bias=False)
        self.decoder.weight = emb_weight
        self.bias = nn.Parameter(torch.zeros(emb_weight.size(0)))


class BertOnlyMLMHead(nn.Module):
    def __init__(self, config, emb_weight):
        super().__init__()
        self.predictions = _BertLMPredictionHead(config, emb_weight)


class GELU(nn.Module):
    pass


class RegionFeatureRegression(nn.Module):
    """model/model.py:1143-1156: weight is the (registered, shared) img_linear weight."""

    def __init__(self, hidden_size, feat_dim, img_linear_weight):
        super().__init__()
        self.net = nn.Sequential(nn.Linear(hidden_size, hidden_size), GELU(), LayerNorm(hidden_size, eps=1e-12))
        self.weight = img_linear_weight
        self.bias = nn.Parameter(torch.zeros(feat_dim))


class RegionClassification(nn.Module):
    def __init__(self, hidden_size, label_dim):
        super().__init__()
        self.net = nn.Sequential(nn.Linear(hidden_size, hidden_size), GELU(), LayerNorm(hidden_size, eps=1e-12),
                                 nn.Linear(hidden_size, label_dim))


def _count(mask, hint):
    return int(hint) if hint is not None else int(mask.sum().item())


# --------------------------------------------------------------------------------------------------
# pretraining model
# --------------------------------------------------------------------------------------------------
class _ForPretraining(UC2PreTrainedModel):
    """MLM + MRFR + MRC + ITM(+WRA OT): model/model.py:460-775 (VLXLMR) / 1172-1501 (Uniter)."""
    encoder_attr = "roberta"
    EncoderCls = VLXLMRModel

    def __init__(self, config, img_dim, img_label_dim, nce_temp=1, ot_pos_only=False):
        super().__init__(config)
        enc = self.EncoderCls(config, img_dim)
        setattr(self, self.encoder_attr, enc)
        W = enc.embeddings.word_embeddings.weight
        if self.family_name == "vlxlmr":
            self.cls = RobertaLMHead(config, W)
        else:
            self.cls = BertOnlyMLMHead(config, W)
        self.feat_regress = RegionFeatureRegression(config.hidden_size, img_dim, enc.img_embeddings.img_linear.weight)
        self.region_classifier = RegionClassification(config.hidden_size, img_label_dim)
        self.itm_output = nn.Linear(config.hidden_size, 2)
        self.ot_pos_only = ot_pos_only
        self.valid_token_ids = None           # VALID_XLMR_TOKEN_IDS of the '*-soft' MRTM tasks (see forward_mmxlm_soft)
        self.apply(self.init_weights)
        self.vocab_pad = 0
        _adopt(self)

    @property
    def _enc(self):
        return getattr(self, self.encoder_attr)

    def pad_vocab(self):
        """No-op like the reference (pad_tensor_to_mul returns early, model/model.py:1051-1054)."""
        self.vocab_pad = 0

    # ---- heads ----------------------------------------------------------------------------------
    def _mlm_transform(self, rows):
        """dense + GELU + LayerNorm of the LM head; returns (h, decoder bias name)."""
        a = self._arena()
        if self.family_name == "vlxlmr":
            h = Fn.LinearFn.apply(rows, a, "cls.dense.weight", "cls.dense.bias", _lib.ACT_GELU, False, False)
            h = Fn.LayerNormFn.apply(h, a, "cls.layer_norm.weight", "cls.layer_norm.bias", float(self.config.layer_norm_eps))
            return h, "cls.bias"
        t = "cls.predictions.transform."
        h = Fn.LinearFn.apply(rows, a, t + "dense.weight", t + "dense.bias", _lib.ACT_GELU, False, False)
        h = Fn.LayerNormFn.apply(h, a, t + "LayerNorm.weight", t + "LayerNorm.bias", 1e-12)
        return h, "cls.predictions.bias"

    def _mlm_scores(self, rows):
        h, bias = self._mlm_transform(rows)
        W = self._enc.prefix + "embeddings.word_embeddings.weight"
        return Fn.LinearFn.apply(h, self._arena(), W, bias, _lib.ACT_NONE, False, True)

    def _mlm_loss(self, rows, labels):
        """Tied decoder + cross entropy fused into one node (no fp32 d(logits), see functional.LmHeadCEFn)."""
        h, bias = self._mlm_transform(rows)
        W = self._enc.prefix + "embeddings.word_embeddings.weight"
        return Fn.LmHeadCEFn.apply(h, self._arena(), W, bias, labels, -100)

    def _feat_regress(self, rows):
        a = self._arena()
        h = Fn.LinearFn.apply(rows, a, "feat_regress.net.0.weight", "feat_regress.net.0.bias", _lib.ACT_GELU, False, False)
        h = Fn.LayerNormFn.apply(h, a, "feat_regress.net.2.weight", "feat_regress.net.2.bias", 1e-12)
        return Fn.LinearFn.apply(h, a, self._enc.prefix + "img_embeddings.img_linear.weight", "feat_regress.bias",
                                 _lib.ACT_NONE, True, True)

    def _region_classify(self, rows):
        a = self._arena()
        h = Fn.LinearFn.apply(rows, a, "region_classifier.net.0.weight", "region_classifier.net.0.bias", _lib.ACT_GELU,
                              False, False)
        h = Fn.LayerNormFn.apply(h, a, "region_classifier.net.2.weight", "region_classifier.net.2.bias", 1e-12)
        return Fn.LinearFn.apply(h, a, "region_classifier.net.3.weight", "region_classifier.net.3.bias", _lib.ACT_NONE,
                                 False, True)

    # ---- dispatch (model/model.py:495-568 / 1206-1265) --------------------------------------------
    def forward(self, batch, task, compute_loss=True):
        batch = defaultdict(lambda: None, batch)
        input_ids = batch["input_ids"]
        if self.family_name == "vlxlmr":
            position_ids = batch["position_ids"] if task == "tlm" else None
        else:
            position_ids = batch["position_ids"]
        img_feat, img_pos_feat = batch["img_feat"], batch["img_pos_feat"]
        attention_mask, gather_index = batch["attn_masks"], batch["gather_index"]
        if task in ("mlm", "tlm"):
            return self.forward_mlm(input_ids, position_ids, img_feat, img_pos_feat, attention_mask, gather_index,
                                    batch["txt_labels"], compute_loss, batch["n_masked"])
        elif task == "tlm-ni":
            return self.forward_mlm(input_ids, position_ids, None, None, attention_mask, None, batch["txt_labels"],
                                    compute_loss, batch["n_masked"])
        elif task == "mrfr":
            return self.forward_mrfr(input_ids, position_ids, img_feat, img_pos_feat, attention_mask, gather_index,
                                     batch["img_masks"], batch["img_mask_tgt"], batch["feat_targets"], compute_loss)
        elif task == "itm":
            return self.forward_itm(input_ids, position_ids, img_feat, img_pos_feat, attention_mask, gather_index,
                                    batch["targets"], batch["ot_inputs"], compute_loss)
        elif task.startswith("mrc"):
            return self.forward_mrc(input_ids, position_ids, img_feat, img_pos_feat, attention_mask, gather_index,
                                    batch["img_masks"], batch["img_mask_tgt"], batch["label_targets"], task, compute_loss)
        elif task in ("mmxlm", "vmlm"):
            return self.forward_mmxlm(input_ids, position_ids, img_feat, img_pos_feat, attention_mask, gather_index,
                                      batch["img_masks"], batch["txt_labels"], compute_loss, batch["n_masked"])
        elif task in ("mmxlm-soft", "vmlm-soft"):
            return self.forward_mmxlm_soft(input_ids, position_ids, img_feat, img_pos_feat, attention_mask,
                                           gather_index, batch["img_masks"], batch["tgt_masks"],
                                           batch["label_targets"], compute_loss)
        else:
            raise ValueError("invalid task")

    def forward_mlm(self, input_ids, position_ids, img_feat, img_pos_feat, attention_mask, gather_index, txt_labels,
                    compute_loss=True, n_masked=None):
        seq = self._enc(input_ids, position_ids, img_feat, img_pos_feat, attention_mask, gather_index,
                        output_all_encoded_layers=False)
        mask = txt_labels != -1                       # text part only: mask columns < T (model.py:583)
        rows = Fn.MaskedRowsFn.apply(seq, mask, _count(mask, n_masked))
        if compute_loss:
            return self._mlm_loss(rows, txt_labels[mask])
        return self._mlm_scores(rows)

    def forward_mmxlm(self, input_ids, position_ids, img_feat, img_pos_feat, attention_mask, gather_index, img_masks,
                      txt_labels, compute_loss=True, n_masked=None):
        """MRTM with hard token labels (model/model.py:598-624): like MLM, but regions are masked too and
        txt_labels covers the whole packed sequence."""
        seq = self._enc(input_ids, position_ids, img_feat, img_pos_feat, attention_mask, gather_index,
                        output_all_encoded_layers=False, img_masks=img_masks)
        mask = txt_labels != -1
        rows = Fn.MaskedRowsFn.apply(seq, mask, _count(mask, n_masked))
        if compute_loss:
            return self._mlm_loss(rows, txt_labels[mask])
        return self._mlm_scores(rows)

    def forward_mmxlm_soft(self, input_ids, position_ids, img_feat, img_pos_feat, attention_mask, gather_index,
                           img_masks, tgt_masks, label_targets, compute_loss=True):
        """MRTM with soft token labels (model/model.py:626-651): vocabulary logits of the masked regions restricted
        to `self.valid_token_ids` (the reference's VALID_XLMR_TOKEN_IDS, model/const_variable.py -- built from the
        XLM-R tokenizer there; here a LongTensor / list the caller assigns), KL against token distributions."""
        if getattr(self, "valid_token_ids", None) is None:
            raise ValueError("set model.valid_token_ids (the reference's VALID_XLMR_TOKEN_IDS) before running "
                             "the '*-soft' MRTM tasks")
        seq = self._enc(input_ids, position_ids, img_feat, img_pos_feat, attention_mask, gather_index,
                        output_all_encoded_layers=False, img_masks=img_masks)
        rows = Fn.MaskedRowsFn.apply(seq, tgt_masks, int(label_targets.size(0)))
        scores = self._mlm_scores(rows)
        cols = torch.as_tensor(self.valid_token_ids, dtype=torch.long, device=scores.device)
        pred = Fn.SelectColumnsFn.apply(scores, cols)
        if compute_loss:
            return Fn.SoftmaxLossFn.apply(pred, 1, label_targets, -1)
        return pred

    def forward_mrfr(self, input_ids, position_ids, img_feat, img_pos_feat, attention_mask, gather_index, img_masks,
                     img_mask_tgt, feat_targets, compute_loss=True):
        seq = self._enc(input_ids, position_ids, img_feat, img_pos_feat, attention_mask, gather_index,
                        output_all_encoded_layers=False, img_masks=img_masks)
        rows = Fn.MaskedRowsFn.apply(seq, img_mask_tgt, int(feat_targets.size(0)))
        pred = self._feat_regress(rows)
        if compute_loss:
            return Fn.MseFn.apply(pred, feat_targets)
        return pred

    def forward_mrc(self, input_ids, position_ids, img_feat, img_pos_feat, attention_mask, gather_index, img_masks,
                    img_mask_tgt, label_targets, task, compute_loss=True):
        seq = self._enc(input_ids, position_ids, img_feat, img_pos_feat, attention_mask, gather_index,
                        output_all_encoded_layers=False, img_masks=img_masks)
        rows = Fn.MaskedRowsFn.apply(seq, img_mask_tgt, int(label_targets.size(0)))
        pred = self._region_classify(rows)
        if not compute_loss:
            return pred
        if "kl" in task:
            return Fn.SoftmaxLossFn.apply(pred, 1, label_targets, -1)
        tgt = torch.max(label_targets[:, 1:], dim=-1)[1] + 1        # background class is never the target
        return Fn.SoftmaxLossFn.apply(pred, 0, tgt, 0)

    def forward_itm(self, input_ids, position_ids, img_feat, img_pos_feat, attention_mask, gather_index, targets,
                    ot_inputs, compute_loss=True):
        seq = self._enc(input_ids, position_ids, img_feat, img_pos_feat, attention_mask, gather_index,
                        output_all_encoded_layers=False)
        pooled = self._enc.pooler(seq)
        rank_scores = Fn.NarrowLinearFn.apply(pooled, self._arena(), "itm_output.weight", "itm_output.bias")
        if ot_inputs is not None:
            from .ot import optimal_transport_dist_packed
            ot_dist = optimal_transport_dist_packed(seq, ot_inputs, input_ids.size(1), img_feat.size(1))
            if self.ot_pos_only:
                ot_loss = ot_dist.masked_select(targets == 1)
            else:
                ot_loss = (ot_dist.masked_select(targets == 1), ot_dist.masked_select(targets == 0))
        else:
            ot_loss = None
        if compute_loss:
            return Fn.SoftmaxLossFn.apply(rank_scores, 0, targets, -100), ot_loss
        return rank_scores, ot_loss


class VLXLMRForPretraining(_ForPretraining):
    family_name = "vlxlmr"
    encoder_attr = "roberta"
    EncoderCls = VLXLMRModel


class UniterForPretraining(_ForPretraining):
    family_name = "uniter"
    encoder_attr = "bert"
    EncoderCls = UniterModel
