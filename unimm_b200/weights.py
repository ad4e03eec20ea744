"""Checkpoint key layout of the reference model and deterministic synthetic weights.

The key names / shapes below are the 535-entry ``state_dict()`` of the reference's
``BertForMultiModalPreTraining`` (reference models/vilbert_dialog.py:1496-1508, module tree
:300-324, :359-612, :615-783, :940-1088, :1475-1485); ``VisualDialogEncoder`` stores it under the
``bert_pretrained.`` prefix (reference models/visual_dialog_encoder.py:14).  Parameters that exist in
checkpoints but never enter the forward (``sep_embeddings``, ``q_dense1/2``) are kept so that a
reference checkpoint loads with ``strict=True``.

``random_state_dict`` draws every tensor from its own ``torch.Generator`` seeded by (seed, key
index), so any process (this repo's tests, the golden-vector script that feeds the *reference*
model, the bench) reproduces bit-identical weights without shipping a 1 GB file.
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Dict, Tuple

import torch

from .config import ViLBertConfig

PREFIX = "bert_pretrained."


def _linear(d: "OrderedDict[str, Tuple[int, ...]]", name: str, out_f: int, in_f: int) -> None:
    d[name + ".weight"] = (out_f, in_f)
    d[name + ".bias"] = (out_f,)


def _ln(d, name: str, n: int) -> None:
    d[name + ".weight"] = (n,)
    d[name + ".bias"] = (n,)


def param_shapes(cfg: ViLBertConfig) -> "OrderedDict[str, Tuple[int, ...]]":
    """Ordered {key: shape} for the un-prefixed reference state dict."""
    H, Hv, Hb = cfg.hidden_size, cfg.v_hidden_size, cfg.bi_hidden_size
    I, Iv = cfg.intermediate_size, cfg.v_intermediate_size
    d: "OrderedDict[str, Tuple[int, ...]]" = OrderedDict()
    e = "bert.embeddings."
    d[e + "word_embeddings.weight"] = (cfg.vocab_size, H)
    d[e + "position_embeddings.weight"] = (cfg.max_position_embeddings, H)
    d[e + "token_type_embeddings.weight"] = (cfg.type_vocab_size, H)
    d[e + "token_type_embeddings_extension.weight"] = (cfg.type_ext_size, H)
    d[e + "sep_embeddings.weight"] = (cfg.sep_embed_size, H)
    _ln(d, e + "LayerNorm", H)
    v = "bert.v_embeddings."
    _linear(d, v + "image_embeddings", Hv, cfg.v_feature_size)
    _linear(d, v + "image_location_embeddings", Hv, cfg.loc_size)
    _ln(d, v + "LayerNorm", Hv)
    for i in range(cfg.num_hidden_layers):
        p = f"bert.encoder.layer.{i}."
        for n in ("query", "key", "value"):
            _linear(d, p + "attention.self." + n, H, H)
        _linear(d, p + "attention.output.dense", H, H)
        _ln(d, p + "attention.output.LayerNorm", H)
        _linear(d, p + "intermediate.dense", I, H)
        _linear(d, p + "output.dense", H, I)
        _ln(d, p + "output.LayerNorm", H)
    for i in range(cfg.v_num_hidden_layers):
        p = f"bert.encoder.v_layer.{i}."
        for n in ("query", "key", "value"):
            _linear(d, p + "attention.self." + n, Hv, Hv)
        _linear(d, p + "attention.output.dense", Hv, Hv)
        _ln(d, p + "attention.output.LayerNorm", Hv)
        _linear(d, p + "intermediate.dense", Iv, Hv)
        _linear(d, p + "output.dense", Hv, Iv)
        _ln(d, p + "output.LayerNorm", Hv)
    for i in range(cfg.num_connections):
        p = f"bert.encoder.c_layer.{i}."
        for n in ("query1", "key1", "value1"):
            _linear(d, p + "biattention." + n, Hb, Hv)
        for n in ("query2", "key2", "value2"):
            _linear(d, p + "biattention." + n, Hb, H)
        _linear(d, p + "biOutput.dense1", Hv, Hb)
        _ln(d, p + "biOutput.LayerNorm1", Hv)
        _linear(d, p + "biOutput.q_dense1", Hv, Hb)
        _linear(d, p + "biOutput.dense2", H, Hb)
        _ln(d, p + "biOutput.LayerNorm2", H)
        _linear(d, p + "biOutput.q_dense2", H, Hb)
        _linear(d, p + "v_intermediate.dense", Iv, Hv)
        _linear(d, p + "v_output.dense", Hv, Iv)
        _ln(d, p + "v_output.LayerNorm", Hv)
        _linear(d, p + "t_intermediate.dense", I, H)
        _linear(d, p + "t_output.dense", H, I)
        _ln(d, p + "t_output.LayerNorm", H)
    _linear(d, "bert.t_pooler.dense", Hb, H)
    _linear(d, "bert.v_pooler.dense", Hb, Hv)
    d["cls.predictions.bias"] = (cfg.vocab_size,)
    _linear(d, "cls.predictions.transform.dense", H, H)
    _ln(d, "cls.predictions.transform.LayerNorm", H)
    d["cls.predictions.decoder.weight"] = (cfg.vocab_size, H)  # tied to word_embeddings (ref :1020)
    _linear(d, "cls.bi_seq_relationship", 2, Hb)
    _linear(d, "cls.imagePredictions.transform.dense", Hv, Hv)
    _ln(d, "cls.imagePredictions.transform.LayerNorm", Hv)
    _linear(d, "cls.imagePredictions.decoder", cfg.v_target_size, Hv)
    return d


TIED = {"cls.predictions.decoder.weight": "bert.embeddings.word_embeddings.weight"}


def random_state_dict(cfg: ViLBertConfig, seed: int = 0, perturbed: bool = False,
                      prefix: str = "", dtype=torch.float32) -> Dict[str, torch.Tensor]:
    """Deterministic synthetic weights in the reference's key layout.

    ``perturbed=False`` follows the reference initialiser's *distribution*
    (``init_bert_weights``, reference models/vilbert_dialog.py:1110-1121): Linear/Embedding weights
    ~ N(0, 0.02), every bias 0, LayerNorm gamma 1 / beta 0.  ``perturbed=True`` additionally draws
    biases ~ N(0, 0.02), gamma ~ N(1, 0.1), beta ~ N(0, 0.1) so that a kernel that drops a bias or
    an affine term cannot pass parity.
    """
    std = cfg.initializer_range
    out: Dict[str, torch.Tensor] = OrderedDict()
    for idx, (name, shape) in enumerate(param_shapes(cfg).items()):
        if name in TIED:
            out[prefix + name] = out[prefix + TIED[name]]
            continue
        g = torch.Generator(device="cpu")
        g.manual_seed((seed * 1_000_003 + idx * 7919 + 12345) & 0x7FFFFFFF)
        is_ln = "LayerNorm" in name
        if name.endswith(".bias"):
            if perturbed:
                t = torch.randn(shape, generator=g, dtype=torch.float32) * (0.1 if is_ln else std)
            else:
                t = torch.zeros(shape, dtype=torch.float32)
        elif is_ln:  # LayerNorm gamma
            t = torch.ones(shape, dtype=torch.float32)
            if perturbed:
                t = t + 0.1 * torch.randn(shape, generator=g, dtype=torch.float32)
        else:
            t = torch.randn(shape, generator=g, dtype=torch.float32) * std
        out[prefix + name] = t.to(dtype)
    return out


def strip_prefix(sd: Dict[str, torch.Tensor], prefix: str = PREFIX) -> Dict[str, torch.Tensor]:
    """Accept either the ``bert_pretrained.``-prefixed or the bare key layout."""
    if any(k.startswith(prefix) for k in sd):
        return OrderedDict((k[len(prefix):], v) for k, v in sd.items() if k.startswith(prefix))
    return sd
