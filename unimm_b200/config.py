"""Model hyper-parameters for the ViLBERT two-stream encoder on the scoring hot path.

Mirrors the fields the reference reads from ``config/bert_base_6layer_6conect.json`` through
``BertConfig.from_json_file`` (reference models/vilbert_dialog.py:249-262).  Fields the JSON omits
take the reference's constructor defaults (models/vilbert_dialog.py:162-169): ``fusion_method='mul'``,
``with_coattention=True``, ``fast_mode=False``, ``fixed_*_layer=0``, ``in_batch_pairs=False``,
``predict_feature=False``.  Only those default values are supported by the CUDA path; anything else
raises at construction time instead of silently computing something different.
"""
from __future__ import annotations

import json
import os
from dataclasses import dataclass, field
from typing import List

DEFAULT_CONFIG_PATH = os.path.join(os.path.dirname(__file__), "config", "bert_base_6layer_6conect.json")


@dataclass
class ViLBertConfig:
    vocab_size: int = 30522
    hidden_size: int = 768
    num_hidden_layers: int = 12
    num_attention_heads: int = 12
    intermediate_size: int = 3072
    max_position_embeddings: int = 512
    type_vocab_size: int = 2
    v_feature_size: int = 2048
    v_target_size: int = 1601
    v_hidden_size: int = 1024
    v_num_hidden_layers: int = 6
    v_num_attention_heads: int = 8
    v_intermediate_size: int = 1024
    bi_hidden_size: int = 1024
    bi_num_attention_heads: int = 8
    v_biattention_id: List[int] = field(default_factory=lambda: [0, 1, 2, 3, 4, 5])
    t_biattention_id: List[int] = field(default_factory=lambda: [6, 7, 8, 9, 10, 11])
    hidden_act: str = "gelu"
    v_hidden_act: str = "gelu"
    initializer_range: float = 0.02
    # reference constructor defaults that the JSON never overrides
    fusion_method: str = "mul"
    with_coattention: bool = True
    fast_mode: bool = False
    fixed_v_layer: int = 0
    fixed_t_layer: int = 0
    in_batch_pairs: bool = False
    predict_feature: bool = False

    # sizes fixed by the reference's embedding module (models/vilbert_dialog.py:317-319)
    type_ext_size: int = 10
    sep_embed_size: int = 50
    loc_size: int = 5

    @classmethod
    def from_dict(cls, d: dict) -> "ViLBertConfig":
        known = {f for f in cls.__dataclass_fields__}
        cfg = cls(**{k: v for k, v in d.items() if k in known})
        cfg.validate()
        return cfg

    @classmethod
    def from_json_file(cls, path: str) -> "ViLBertConfig":
        with open(path, "r", encoding="utf-8") as f:
            return cls.from_dict(json.load(f))

    def validate(self) -> None:
        if self.hidden_act != "gelu" or self.v_hidden_act != "gelu":
            raise ValueError("only the erf-GELU activation of the reference config is supported")
        if self.fusion_method != "mul" or not self.with_coattention or self.fast_mode \
                or self.in_batch_pairs or self.predict_feature or self.fixed_t_layer or self.fixed_v_layer:
            raise ValueError("unsupported ViLBERT variant: only the reference defaults are implemented")
        if len(self.v_biattention_id) != len(self.t_biattention_id):
            raise ValueError("v_biattention_id and t_biattention_id must have equal length")
        if max(self.v_biattention_id) >= self.v_num_hidden_layers or max(self.t_biattention_id) >= self.num_hidden_layers:
            raise ValueError("bi-attention ids out of range")
        if self.hidden_size % self.num_attention_heads or self.v_hidden_size % self.v_num_attention_heads \
                or self.bi_hidden_size % self.bi_num_attention_heads:
            raise ValueError("hidden sizes must be divisible by head counts")

    @property
    def head_dim(self) -> int:
        return self.hidden_size // self.num_attention_heads

    @property
    def v_head_dim(self) -> int:
        return self.v_hidden_size // self.v_num_attention_heads

    @property
    def bi_head_dim(self) -> int:
        return self.bi_hidden_size // self.bi_num_attention_heads

    @property
    def num_connections(self) -> int:
        return len(self.v_biattention_id)

    def layer_schedule(self):
        """The order in which BertEncoder.forward runs layers (reference models/vilbert_dialog.py:842-929).

        Returns a list of ("t", i) / ("v", i) / ("c", i) tuples.
        """
        sched = []
        v_start = t_start = 0
        for c, (v_end, t_end) in enumerate(zip(self.v_biattention_id, self.t_biattention_id)):
            sched += [("v", i) for i in range(v_start, v_end)]
            sched += [("t", i) for i in range(t_start, t_end)]
            sched.append(("c", c))
            v_start, t_start = v_end, t_end
        sched += [("v", i) for i in range(v_start, self.v_num_hidden_layers)]
        sched += [("t", i) for i in range(t_start, self.num_hidden_layers)]
        return sched


def tiny_config(**kw) -> ViLBertConfig:
    """Same widths as the reference config, fewer layers: for fast CPU-side tests."""
    base = dict(num_hidden_layers=2, v_num_hidden_layers=1, v_biattention_id=[0], t_biattention_id=[1])
    base.update(kw)
    return ViLBertConfig.from_dict(base)
