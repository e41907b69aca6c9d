"""BERT-family model description + weight sources for the B200 encoder.

The reference names its models by HuggingFace id (``embedding_model`` /
``rerank_model`` in ``server/.env.template:3,43``; constructed at
``server/RAGHelper_local.py:114-117`` and ``server/RAGHelper.py:484``).  This
module resolves such a name to (config, weight dict, pooling rule) from

* a local HuggingFace / sentence-transformers snapshot directory
  (``config.json`` + ``model.safetensors`` [+ ``modules.json``,
  ``1_Pooling/config.json``, ``sentence_bert_config.json``]); or
* ``synthetic:<preset>[:seed]`` — seeded random-init weights of the named
  architecture.  There is no network in the build/bench image, so the
  benchmarks and parity tests use these (BASELINE.json configs only fix the
  *shapes*).

Weight dict keys are HuggingFace ``BertModel`` names without the ``bert.``
prefix; values are contiguous float32 numpy arrays.
"""
from __future__ import annotations

import json
import os
from dataclasses import dataclass, asdict
from typing import Dict, Optional, Tuple

import numpy as np


@dataclass
class BertConfig:
    vocab_size: int = 30522
    hidden: int = 384
    layers: int = 6
    heads: int = 12
    ffn: int = 1536
    max_pos: int = 512
    type_vocab: int = 2
    ln_eps: float = 1e-12
    num_labels: int = 1

    @property
    def head_dim(self) -> int:
        return self.hidden // self.heads

    @staticmethod
    def from_hf(d: dict) -> "BertConfig":
        if d.get("hidden_act", "gelu") != "gelu":
            raise ValueError(f"only erf-GELU BERT encoders are supported, got {d.get('hidden_act')}")
        if d.get("position_embedding_type", "absolute") != "absolute":
            raise ValueError("only absolute position embeddings are supported")
        return BertConfig(
            vocab_size=d["vocab_size"], hidden=d["hidden_size"], layers=d["num_hidden_layers"],
            heads=d["num_attention_heads"], ffn=d["intermediate_size"],
            max_pos=d["max_position_embeddings"], type_vocab=d.get("type_vocab_size", 2),
            ln_eps=d.get("layer_norm_eps", 1e-12),
            num_labels=len(d["id2label"]) if "id2label" in d else d.get("num_labels", 1))


# shapes of the models BASELINE.json / .env.template name
PRESETS: Dict[str, Tuple[BertConfig, str, int]] = {
    # name: (config, pooling, max_seq_length)
    "all-MiniLM-L6-v2": (BertConfig(hidden=384, layers=6, heads=12, ffn=1536), "mean", 256),
    "ms-marco-MiniLM-L-6-v2": (BertConfig(hidden=384, layers=6, heads=12, ffn=1536, num_labels=1), "cls", 512),
    "bge-base-en-v1.5": (BertConfig(hidden=768, layers=12, heads=12, ffn=3072), "cls", 512),
    "GIST-small-Embedding-v0": (BertConfig(hidden=384, layers=12, heads=12, ffn=1536), "cls", 512),
    # tiny shapes for unit tests
    "tiny": (BertConfig(vocab_size=1000, hidden=128, layers=2, heads=4, ffn=256, max_pos=128), "mean", 64),
}


def weight_names(cfg: BertConfig, with_head: bool):
    H, F = cfg.hidden, cfg.ffn
    names = [
        ("embeddings.word_embeddings.weight", (cfg.vocab_size, H)),
        ("embeddings.position_embeddings.weight", (cfg.max_pos, H)),
        ("embeddings.token_type_embeddings.weight", (cfg.type_vocab, H)),
        ("embeddings.LayerNorm.weight", (H,)),
        ("embeddings.LayerNorm.bias", (H,)),
    ]
    for l in range(cfg.layers):
        p = f"encoder.layer.{l}."
        names += [
            (p + "attention.self.query.weight", (H, H)), (p + "attention.self.query.bias", (H,)),
            (p + "attention.self.key.weight", (H, H)), (p + "attention.self.key.bias", (H,)),
            (p + "attention.self.value.weight", (H, H)), (p + "attention.self.value.bias", (H,)),
            (p + "attention.output.dense.weight", (H, H)), (p + "attention.output.dense.bias", (H,)),
            (p + "attention.output.LayerNorm.weight", (H,)), (p + "attention.output.LayerNorm.bias", (H,)),
            (p + "intermediate.dense.weight", (F, H)), (p + "intermediate.dense.bias", (F,)),
            (p + "output.dense.weight", (H, F)), (p + "output.dense.bias", (H,)),
            (p + "output.LayerNorm.weight", (H,)), (p + "output.LayerNorm.bias", (H,)),
        ]
    if with_head:
        names += [
            ("pooler.dense.weight", (H, H)), ("pooler.dense.bias", (H,)),
            ("classifier.weight", (cfg.num_labels, H)), ("classifier.bias", (cfg.num_labels,)),
        ]
    return names


def synthetic_bert_weights(cfg: BertConfig, seed: int = 0, with_head: bool = False,
                           scale: float = 1.0) -> Dict[str, np.ndarray]:
    """Seeded random weights.  ``scale=1`` is BERT's init (N(0, 0.02)) with small
    random biases and LayerNorm gains near 1; ``scale>1`` widens the linear
    weights to a "trained-like" regime (saturating softmax, logits of order 10)."""
    rng = np.random.default_rng(seed)
    out: Dict[str, np.ndarray] = {}
    for name, shape in weight_names(cfg, with_head):
        if name.endswith("LayerNorm.weight"):
            a = 1.0 + 0.1 * rng.standard_normal(shape)
        elif name.endswith("LayerNorm.bias"):
            a = 0.05 * rng.standard_normal(shape)
        elif name.endswith(".bias"):
            a = 0.02 * scale * rng.standard_normal(shape)
        elif name.startswith("embeddings."):
            a = 0.02 * rng.standard_normal(shape) * (3.0 if scale > 1 else 1.0)
        else:
            a = 0.02 * scale * rng.standard_normal(shape)
        out[name] = np.ascontiguousarray(a, dtype=np.float32)
    return out


def _read_json(path: str) -> Optional[dict]:
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f)
    return None


def load_hf_snapshot(path: str, with_head: bool):
    """Read a local HF / sentence-transformers snapshot.  Returns
    (config, weights, pooling, normalize, max_seq_length, activation)."""
    from safetensors.numpy import load_file
    cfgd = _read_json(os.path.join(path, "config.json"))
    if cfgd is None:
        raise FileNotFoundError(f"{path}: no config.json")
    cfg = BertConfig.from_hf(cfgd)
    raw = load_file(os.path.join(path, "model.safetensors"))
    w: Dict[str, np.ndarray] = {}
    for k, v in raw.items():
        k2 = k[5:] if k.startswith("bert.") else k
        w[k2] = np.ascontiguousarray(v, dtype=np.float32)
    need = [n for n, _ in weight_names(cfg, with_head)]
    missing = [n for n in need if n not in w]
    if missing:
        raise KeyError(f"{path}: missing tensors {missing[:4]}...")
    pooling, normalize, max_len = "mean", False, min(cfg.max_pos, 512)
    mods = _read_json(os.path.join(path, "modules.json")) or []
    for m in mods:
        if m.get("type", "").endswith("Pooling"):
            pc = _read_json(os.path.join(path, m["path"], "config.json")) or {}
            if pc.get("pooling_mode_cls_token"):
                pooling = "cls"
            elif pc.get("pooling_mode_mean_tokens", True):
                pooling = "mean"
            else:
                raise ValueError("unsupported pooling mode in " + m["path"])
        if m.get("type", "").endswith("Normalize"):
            normalize = True
    sb = _read_json(os.path.join(path, "sentence_bert_config.json"))
    if sb and sb.get("max_seq_length"):
        max_len = int(sb["max_seq_length"])
    act = "identity"
    a = cfgd.get("sbert_ce_default_activation_function")
    if with_head:
        if a is None:
            act = "sigmoid" if cfg.num_labels == 1 else "identity"
        else:
            act = "sigmoid" if "Sigmoid" in a else "identity"
    return cfg, w, pooling, normalize, max_len, act


def resolve_model(model_name: str, with_head: bool, cache_folder: Optional[str] = None):
    """model id -> (cfg, weights, pooling, normalize, max_seq_length, activation, vocab_source).

    ``synthetic:<preset>[:seed[:scale]]`` builds seeded weights; anything else must be a
    local directory (or a directory ``<cache>/<name>`` / ``$RAGMEUP_MODEL_DIR/<name>``)."""
    if model_name.startswith("synthetic:"):
        parts = model_name.split(":")
        preset = parts[1]
        seed = int(parts[2]) if len(parts) > 2 else 0
        scale = float(parts[3]) if len(parts) > 3 else 1.0
        if preset not in PRESETS:
            raise ValueError(f"unknown synthetic preset {preset!r}; have {sorted(PRESETS)}")
        cfg, pooling, max_len = PRESETS[preset]
        cfg = BertConfig(**asdict(cfg))
        w = synthetic_bert_weights(cfg, seed=seed, with_head=with_head, scale=scale)
        return cfg, w, pooling, True, max_len, "identity", None
    cands = [model_name]
    for root in (cache_folder, os.environ.get("RAGMEUP_MODEL_DIR")):
        if root:
            cands += [os.path.join(root, model_name), os.path.join(root, model_name.split("/")[-1])]
    for c in cands:
        if os.path.isdir(c):
            return (*load_hf_snapshot(c, with_head), c)
    raise FileNotFoundError(
        f"model {model_name!r} not found locally (no network in this deployment): point "
        f"RAGMEUP_MODEL_DIR at a directory holding its HuggingFace snapshot, or use 'synthetic:<preset>'")
