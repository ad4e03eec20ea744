"""Python handle on one ``unimm_engine_t`` (one per CUDA device): weight upload, forward, host-buffer scoring.

PyTorch is used only for device memory, streams and dtype plumbing; all arithmetic happens in
``libunimm_b200.so`` (see include/unimm_b200.h).
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional

import torch

from . import _lib
from ._lib import Batch, Config, HostBatch, Outputs, check, lib, ptr
from .config import ViLBertConfig
from .weights import strip_prefix

PRECISIONS = {"fp32": _lib.PREC_FP32, "bf16": _lib.PREC_BF16, "fp16": _lib.PREC_FP16}


def c_config(cfg: ViLBertConfig, seq_len: int = 256, num_regions: int = 37) -> Config:
    c = Config()
    for name in ("vocab_size", "hidden_size", "num_hidden_layers", "num_attention_heads", "intermediate_size",
                 "max_position_embeddings", "type_vocab_size", "v_feature_size", "v_target_size", "v_hidden_size",
                 "v_num_hidden_layers", "v_num_attention_heads", "v_intermediate_size", "bi_hidden_size",
                 "bi_num_attention_heads"):
        setattr(c, name, int(getattr(cfg, name)))
    c.num_connections = cfg.num_connections
    for i, (v, t) in enumerate(zip(cfg.v_biattention_id, cfg.t_biattention_id)):
        c.v_biattention_id[i] = v
        c.t_biattention_id[i] = t
    c.seq_len = seq_len
    c.num_regions = num_regions
    return c


def _i64(t: torch.Tensor, device) -> torch.Tensor:
    return t.to(device=device, dtype=torch.int64, non_blocking=True).contiguous()


def _f32(t: torch.Tensor, device) -> torch.Tensor:
    return t.to(device=device, dtype=torch.float32, non_blocking=True).contiguous()


class Engine:
    """Weights + workspace on one GPU.  ``max_sequences`` bounds the chunk size of one forward."""

    def __init__(self, cfg: ViLBertConfig, state_dict: Dict[str, torch.Tensor], precision: str = "bf16",
                 max_sequences: int = 128, device: Optional[int] = None, seq_len: int = 256, num_regions: int = 37):
        if not torch.cuda.is_available():
            raise RuntimeError("unimm_b200 needs a CUDA device (B200 / sm_100a); there is no CPU path")
        if precision not in PRECISIONS:
            raise ValueError(f"precision must be one of {sorted(PRECISIONS)}")
        self.cfg, self.precision, self.max_sequences = cfg, precision, int(max_sequences)
        self.seq_len, self.num_regions = seq_len, num_regions
        self.device_index = torch.cuda.current_device() if device is None else int(device)
        self.device = torch.device("cuda", self.device_index)
        self._h = C.c_void_p()
        ccfg = c_config(cfg, seq_len, num_regions)
        check(lib.unimm_create(C.byref(ccfg), self.device_index, PRECISIONS[precision], self.max_sequences, C.byref(self._h)))
        try:
            for name, t in strip_prefix(state_dict).items():
                t = t.detach().to("cpu", torch.float32).contiguous()
                shape = (C.c_int64 * t.dim())(*t.shape)
                check(lib.unimm_load_weight(self._h, name.encode(), C.c_void_p(t.data_ptr()), shape, t.dim()))
            check(lib.unimm_finalize_weights(self._h))
        except Exception:
            self.close()
            raise

    def close(self) -> None:
        if getattr(self, "_h", None) is not None and self._h.value:
            lib.unimm_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------------------------------ forward
    def forward(self, input_ids, token_type_ids, position_ids, desc, image_feat, image_loc, image_mask,
                feat_index=None, masked_lm_labels=None, lm_rows=None, lm_weight=None, next_sentence_label=None,
                image_label=None, image_target=None, nsp_weight=None, want=("seq_score", "nsp_scores")) -> Dict[str, torch.Tensor]:
        """One chunk (B <= max_sequences).  All tensors are moved to this engine's device if needed.

        ``desc`` is int32 [B,4] (mode, ctx, L, last_len).  ``want`` selects outputs among
        seq_score, token_logp, token_ul, nsp_scores, losses, sequence_output_t, sequence_output_v,
        prediction_scores_t.
        """
        dev = self.device
        B, S, R = int(input_ids.shape[0]), self.seq_len, self.num_regions
        if B > self.max_sequences:
            raise ValueError(f"chunk of {B} sequences exceeds max_sequences={self.max_sequences}")
        if tuple(input_ids.shape) != (B, S):
            raise ValueError(f"input_ids must be [B,{S}]")
        keep = []  # keep device tensors alive until the call returns

        def hold(t):
            keep.append(t)
            return t

        b = Batch()
        b.B = B
        b.d_input_ids = ptr(hold(_i64(input_ids, dev)))
        b.d_token_type_ids = ptr(hold(_i64(token_type_ids, dev)))
        b.d_position_ids = ptr(hold(_i64(position_ids, dev)))
        d = hold(desc.to(device=dev, dtype=torch.int32).contiguous())
        if tuple(d.shape) != (B, 4):
            raise ValueError("desc must be int32 [B,4]")
        b.d_desc = ptr(d)
        feat = hold(_f32(image_feat, dev))
        U = int(feat.shape[0])
        if tuple(feat.shape[1:]) != (R, self.cfg.v_feature_size):
            raise ValueError("image_feat must be [U,R,v_feature_size]")
        b.d_image_feat = ptr(feat)
        b.d_image_loc = ptr(hold(_f32(image_loc, dev)))
        b.d_image_mask = ptr(hold(_f32(image_mask, dev)))
        if feat_index is not None:
            b.d_feat_index = ptr(hold(feat_index.to(device=dev, dtype=torch.int32).contiguous()))
        elif U != B:
            raise ValueError("feat_index is required when image tensors are not per-sequence")
        n_rows = 0
        if masked_lm_labels is not None:
            labels = hold(_i64(masked_lm_labels, dev))
            b.d_masked_lm_labels = ptr(labels)
            if lm_rows is None:
                lm_rows = (labels.view(-1) != -1).nonzero().view(-1)
            rows = hold(lm_rows.to(device=dev, dtype=torch.int32).contiguous())
            n_rows = int(rows.numel())
            b.d_lm_rows = ptr(rows)
        b.n_lm_rows = n_rows
        if lm_weight is not None:
            b.d_lm_weight = ptr(hold(_i64(lm_weight, dev)))
        if next_sentence_label is not None:
            b.d_next_sentence_label = ptr(hold(_i64(next_sentence_label, dev)))
        if image_label is not None:
            b.d_image_label = ptr(hold(_i64(image_label, dev)))
        if image_target is not None:
            b.d_image_target = ptr(hold(_f32(image_target, dev)))
        if nsp_weight is not None:
            b.d_nsp_weight = ptr(hold(_f32(nsp_weight.reshape(-1)[:2], dev)))

        shapes = {"seq_score": (B,), "token_logp": (B, S), "token_ul": (B, S), "nsp_scores": (B, 2), "losses": (8,),
                  "sequence_output_t": (B, S, self.cfg.hidden_size), "sequence_output_v": (B, R, self.cfg.v_hidden_size),
                  "prediction_scores_t": (B, S, self.cfg.vocab_size)}
        out, o = {}, Outputs()
        for name in want:
            out[name] = torch.zeros(shapes[name], dtype=torch.float32, device=dev)
            setattr(o, "d_" + name, ptr(out[name]))
        stream = torch.cuda.current_stream(dev).cuda_stream
        check(lib.unimm_forward(self._h, C.byref(b), C.byref(o), C.c_void_p(stream)))
        out["_keepalive"] = keep
        return out

    def check_ids(self) -> None:
        """Synchronise the current stream and raise if the last forward met a token / position / token-type id outside its
        embedding table (the reference's nn.Embedding raises an IndexError there; the kernels clamp and flag)."""
        stream = torch.cuda.current_stream(self.device).cuda_stream
        check(lib.unimm_check_ids(self._h, C.c_void_p(stream)))

    # ------------------------------------------------------------------------------------------ prefix-shared path
    def forward_packed(self, pb, want=("seq_score", "nsp_scores")) -> Dict[str, torch.Tensor]:
        """Prefix-shared generative scoring of a ``packing.PackedBatch`` whose tensors are on this device."""
        dev = self.device
        for k, t in pb.tensors().items():
            if t.device != dev:
                raise ValueError(f"packed tensor {k} is on {t.device}, expected {dev} (use PackedBatch.to)")
        out = {}
        if "seq_score" in want:
            out["seq_score"] = torch.zeros(pb.n_cands, device=dev)
        if "nsp_scores" in want:
            out["nsp_scores"] = torch.zeros(pb.n_cands, 2, device=dev)
        if "token_logp" in want:
            out["token_logp"] = torch.zeros(pb.lm_rows.shape[0], device=dev)
        s = pb.c_struct()
        stream = torch.cuda.current_stream(dev).cuda_stream
        check(lib.unimm_forward_packed(self._h, C.byref(s), ptr(out.get("seq_score")), ptr(out.get("nsp_scores")),
                                       ptr(out.get("token_logp")), C.c_void_p(stream)))
        return out

    def score_packed_host(self, pb, seq_score: torch.Tensor, nsp_scores: Optional[torch.Tensor] = None) -> None:
        """End to end from (pinned) host tensors: H2D of the packed arrays + forward + D2H of the scores + sync."""
        for k, t in pb.tensors().items():
            if t.device.type != "cpu":
                raise ValueError(f"packed tensor {k} must be a host tensor")
        s = pb.c_struct()
        stream = torch.cuda.current_stream(self.device).cuda_stream
        check(lib.unimm_score_packed_host(self._h, C.byref(s), ptr(seq_score), ptr(nsp_scores), C.c_void_p(stream)))

    def submit_packed_host(self, pb, slot: int, seq_score: torch.Tensor, nsp_scores: Optional[torch.Tensor] = None) -> None:
        """Asynchronous half of ``score_packed_host``: enqueue H2D + forward + D2H into staging slot 0 / 1 and return; ``pb`` and the
        (pinned) result tensors must stay untouched until ``wait_packed(slot)``."""
        s = pb.c_struct()
        stream = torch.cuda.current_stream(self.device).cuda_stream
        check(lib.unimm_submit_packed_host(self._h, C.byref(s), int(slot), ptr(seq_score), ptr(nsp_scores), C.c_void_p(stream)))

    def wait_packed(self, slot: int) -> None:
        check(lib.unimm_wait_packed(self._h, int(slot)))

    # ------------------------------------------------------------------------------------------ profiling
    PROFILE_CLASSES = ("gemm", "attention", "layernorm", "lm_head", "gemm_ln")     # gemm = umma_gemm_kernel, gemm_ln = umma_gemm_ln_kernel

    def profile_begin(self) -> None:
        check(lib.unimm_profile_begin(self._h))

    def profile_end(self) -> Dict[str, Dict[str, float]]:
        """{class: {ms, work, launches, bytes}}; work = FLOPs (bytes for layernorm), bytes = algorithmic HBM bytes (gemm class); synchronises."""
        n = len(self.PROFILE_CLASSES)
        ms, work, cnt = (C.c_double * n)(), (C.c_double * n)(), (C.c_int64 * n)()
        check(lib.unimm_profile_end(self._h, ms, work, cnt, n))
        nbytes = (C.c_double * n)()
        check(lib.unimm_profile_bytes(self._h, nbytes, n))
        return {k: {"ms": ms[i], "work": work[i], "launches": int(cnt[i]), "bytes": nbytes[i]} for i, k in enumerate(self.PROFILE_CLASSES)}

    # ------------------------------------------------------------------------------------------ host path
    def score_host(self, hb: "HostArrays", seq_score: torch.Tensor, nsp_scores: Optional[torch.Tensor] = None) -> None:
        """End-to-end scoring from (pinned) host arrays: H2D + forward + D2H + sync inside one C call."""
        h = HostBatch()
        h.B, h.U = hb.B, hb.U
        for f in ("input_ids", "token_type_ids", "position_ids", "masked_lm_labels", "desc", "image_feat", "image_loc",
                  "image_mask", "feat_index"):
            setattr(h, "h_" + f, ptr(getattr(hb, f)))
        stream = torch.cuda.current_stream(self.device).cuda_stream
        check(lib.unimm_score_host(self._h, C.byref(h), ptr(seq_score), ptr(nsp_scores), C.c_void_p(stream)))


class HostArrays:
    """Host-side (ideally pinned) inputs of ``Engine.score_host``; shapes follow include/unimm_b200.h."""

    def __init__(self, input_ids, token_type_ids, position_ids, masked_lm_labels, desc, image_feat, image_loc, image_mask,
                 feat_index=None):
        self.B, self.U = int(input_ids.shape[0]), int(image_feat.shape[0])
        for n, t, dt in (("input_ids", input_ids, torch.int64), ("token_type_ids", token_type_ids, torch.int64),
                         ("position_ids", position_ids, torch.int64), ("masked_lm_labels", masked_lm_labels, torch.int64),
                         ("desc", desc, torch.int32), ("image_feat", image_feat, torch.float32),
                         ("image_loc", image_loc, torch.float32), ("image_mask", image_mask, torch.float32),
                         ("feat_index", feat_index, torch.int32)):
            if t is not None:
                assert t.device.type == "cpu" and t.dtype == dt and t.is_contiguous(), n
            setattr(self, n, t)

    def bytes_h2d(self) -> int:
        return sum(t.numel() * t.element_size() for t in (self.input_ids, self.token_type_ids, self.position_ids,
                                                          self.masked_lm_labels, self.desc, self.image_feat, self.image_loc,
                                                          self.image_mask, self.feat_index) if t is not None)
