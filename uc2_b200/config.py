"""Model configuration (mirror of VLXLMRConfig / UniterConfig,
/root/reference/model/model.py:45-143, 781-870) and the state_dict naming contract
(SURVEY.md 8b)."""
import copy
import json

from .synth import IMG_DIM, IMG_LABEL_DIM, UC2_BASE


class UC2Config(object):
    def __init__(self, vocab_size_or_config_json_file=None, **kw):
        d = dict(UC2_BASE)
        if isinstance(vocab_size_or_config_json_file, str):
            with open(vocab_size_or_config_json_file, "r", encoding="utf-8") as f:
                d.update(json.loads(f.read()))
        elif isinstance(vocab_size_or_config_json_file, int):
            d["vocab_size"] = vocab_size_or_config_json_file
        elif vocab_size_or_config_json_file is not None:
            raise ValueError("First argument must be either a vocabulary size (int) or the path "
                             "to a pretrained model config file (str)")
        d.update(kw)
        self.__dict__.update(d)

    @classmethod
    def from_dict(cls, obj):
        c = cls()
        c.__dict__.update(obj)
        return c

    @classmethod
    def from_json_file(cls, path):
        return cls(path)

    def to_dict(self):
        return copy.deepcopy(self.__dict__)

    def to_json_string(self):
        return json.dumps(self.to_dict(), indent=2, sort_keys=True) + "\n"

    def __repr__(self):
        return self.to_json_string()

    def check_kernel_support(self):
        """The CUDA kernels specialise on the uc2-base / bert-base geometry; anything else
        raises (no fallback path exists)."""
        if (self.hidden_size, self.num_attention_heads, self.intermediate_size) != (768, 12, 3072):
            raise ValueError("uc2_b200 kernels are specialised for hidden 768 / 12 heads / 3072 FFN; got "
                             f"{self.hidden_size}/{self.num_attention_heads}/{self.intermediate_size}")
        if self.hidden_act != "gelu":
            raise ValueError("only hidden_act='gelu' (erf form, model/layer.py:31-37) is implemented")
        # packed sequences (text + regions) above 256 run on the tiled mma.sync attention kernels, which implement no
        # attention-probability dropout: say so when the model is built, not at the first training step
        if self.attention_probs_dropout_prob > 0 and getattr(self, "max_packed_len", 256) > 256:
            raise ValueError("attention dropout is implemented for packed lengths up to 256; set max_packed_len <= 256 "
                             "(every UC2 recipe: <= 122 tokens + 100 regions) or attention_probs_dropout_prob = 0")


VLXLMRConfig = UC2Config
UniterConfig = UC2Config


def encoder_shapes(cfg, family="vlxlmr", img_dim=IMG_DIM):
    H, I = cfg.hidden_size, cfg.intermediate_size
    p = "roberta." if family == "vlxlmr" else "bert."
    typ = "new_token_type_embeddings" if family == "vlxlmr" else "token_type_embeddings"
    s = {
        p + "embeddings.word_embeddings.weight": (cfg.vocab_size, H),
        p + "embeddings.position_embeddings.weight": (cfg.max_position_embeddings, H),
        p + f"embeddings.{typ}.weight": (cfg.type_vocab_size, H),
        p + "embeddings.LayerNorm.weight": (H,), p + "embeddings.LayerNorm.bias": (H,),
        p + "img_embeddings.img_linear.weight": (H, img_dim), p + "img_embeddings.img_linear.bias": (H,),
        p + "img_embeddings.img_layer_norm.weight": (H,), p + "img_embeddings.img_layer_norm.bias": (H,),
        p + "img_embeddings.pos_layer_norm.weight": (H,), p + "img_embeddings.pos_layer_norm.bias": (H,),
        p + "img_embeddings.pos_linear.weight": (H, 7), p + "img_embeddings.pos_linear.bias": (H,),
        p + "img_embeddings.mask_embedding.weight": (2, img_dim),
        p + "img_embeddings.LayerNorm.weight": (H,), p + "img_embeddings.LayerNorm.bias": (H,),
    }
    for l in range(cfg.num_hidden_layers):
        q = p + f"encoder.layer.{l}."
        for n in ("query", "key", "value"):
            s[q + f"attention.self.{n}.weight"] = (H, H)
            s[q + f"attention.self.{n}.bias"] = (H,)
        s[q + "attention.output.dense.weight"] = (H, H)
        s[q + "attention.output.dense.bias"] = (H,)
        s[q + "attention.output.LayerNorm.weight"] = (H,)
        s[q + "attention.output.LayerNorm.bias"] = (H,)
        s[q + "intermediate.dense.weight"] = (I, H)
        s[q + "intermediate.dense.bias"] = (I,)
        s[q + "output.dense.weight"] = (H, I)
        s[q + "output.dense.bias"] = (H,)
        s[q + "output.LayerNorm.weight"] = (H,)
        s[q + "output.LayerNorm.bias"] = (H,)
    s[p + "pooler.dense.weight"] = (H, H)
    s[p + "pooler.dense.bias"] = (H,)
    return s


def pretraining_shapes(cfg, family="vlxlmr", img_dim=IMG_DIM, img_label_dim=IMG_LABEL_DIM):
    """Unique parameters of {VLXLMR,Uniter}ForPretraining (tied aliases cls.decoder.weight,
    cls.decoder.bias, feat_regress.weight are NOT listed: they share storage)."""
    H = cfg.hidden_size
    s = encoder_shapes(cfg, family, img_dim)
    if family == "vlxlmr":
        s.update({"cls.bias": (cfg.vocab_size,), "cls.dense.weight": (H, H), "cls.dense.bias": (H,),
                  "cls.layer_norm.weight": (H,), "cls.layer_norm.bias": (H,)})
    else:
        t = "cls.predictions."
        s.update({t + "bias": (cfg.vocab_size,), t + "transform.dense.weight": (H, H),
                  t + "transform.dense.bias": (H,), t + "transform.LayerNorm.weight": (H,),
                  t + "transform.LayerNorm.bias": (H,)})
    s.update({"feat_regress.bias": (img_dim,), "feat_regress.net.0.weight": (H, H),
              "feat_regress.net.0.bias": (H,), "feat_regress.net.2.weight": (H,),
              "feat_regress.net.2.bias": (H,),
              "region_classifier.net.0.weight": (H, H), "region_classifier.net.0.bias": (H,),
              "region_classifier.net.2.weight": (H,), "region_classifier.net.2.bias": (H,),
              "region_classifier.net.3.weight": (img_label_dim, H),
              "region_classifier.net.3.bias": (img_label_dim,),
              "itm_output.weight": (2, H), "itm_output.bias": (2,)})
    return s


def retrieval_shapes(cfg, family="vlxlmr", img_dim=IMG_DIM):
    H = cfg.hidden_size
    s = encoder_shapes(cfg, family, img_dim)
    s.update({"itm_output.weight": (2, H), "itm_output.bias": (2,),
              "rank_output.weight": (1, H), "rank_output.bias": (1,)})
    return s
