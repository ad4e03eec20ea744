"""One training step of the reference on the device: forward, the three losses, the backward of every layer and AdamW
(reference train.py:445-463 — ``forward`` under autocast, ``scaler.scale(loss).backward()``,
``scaler.step(optimizer)`` — with the parameter groups of train.py:322-347).

What replaces what
------------------
* autograd            -> the layer schedule (models/vilbert_dialog.py:842-929) walked in reverse over activations SAVED by the forward
                         below: per projection its 16-bit input, per LayerNorm its fp32 input, per GELU its pre-activation, per
                         attention the 16-bit Q / K / V / context and the row log-sum-exp (no probabilities, no masks).
* autocast + GradScaler -> fp32 master weights, fp32 gradients and accumulators; every GEMM / attention operand is 16-bit.  Gradients
                         that become tensor-core operands get a power-of-two scale from their own maximum ON THE DEVICE
                         (``unimm_k_linear_backward`` / ``unimm_k_attention_backward``), so there is no global loss scale, no
                         inf / nan check and no skipped step.
* pytorch_transformers.AdamW -> ``unimm_t_adamw`` over four contiguous ranges of one flat parameter buffer (language / vision
                         learning rate x decay / no decay, train.py:322-345), which also refreshes the 16-bit operand copies.

Every arithmetic operation is a kernel of ``libunimm_b200.so`` reached through ``ops`` (``unimm_b200.train_ops.DeviceOps``);
``tests/`` substitutes a torch-fp64 statement of the same operations to check this file's orchestration against ``torch.autograd``
of the oracle on the CPU.  Dropout: ``TrainStep(dropout=0.1)`` applies the reference's training-mode ``nn.Dropout`` at every one of its call
sites with counter-based masks that the backward regenerates (``site_seed``, csrc/common.cuh ``drop_keep``); the default 0 is the reference
with its dropout layers in eval mode.

Layout: Q | K | V weights of a layer are adjacent in the flat buffer, so the fused ``[3H, K]`` projection and its gradient are
views; three tensors are stored padded to tensor-core friendly shapes (their padding stays exactly zero under AdamW): the NSP head
``[2, 1024] -> [64, 1024]``, the image-class decoder ``[1601, 1024] -> [1664, 1024]`` and the location projection
``[1024, 5] -> [1024, 64]``.
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Dict, Optional

import zlib

import numpy as np
import torch

from .config import ViLBertConfig
from .train_ops import ACT_GELU, ACT_NONE, ACT_RELU, EW_ADD, EW_SCALE, MASK_CO_INTERVAL, MASK_KEY_VECTOR, MASK_TEXT_SELF
from .weights import param_shapes

NO_DECAY = ("bias", "LayerNorm.bias", "LayerNorm.weight")            # train.py:323 (substring match, so LayerNorm1/2.weight DO decay)
TIED = {"cls.predictions.decoder.weight": "bert.embeddings.word_embeddings.weight"}
UNUSED_MARKERS = ("sep_embeddings", "q_dense1", "q_dense2")          # parameters the forward never touches: grad None in the reference
PADDED = {"cls.bi_seq_relationship.weight": (64, None), "cls.bi_seq_relationship.bias": (64,),
          "cls.imagePredictions.decoder.weight": (1664, None), "cls.imagePredictions.decoder.bias": (1664,),
          "bert.v_embeddings.image_location_embeddings.weight": (None, 64)}


def is_language_weight(name: str) -> bool:
    """Membership in the reference's config/language_weights.json for the names the model actually has (train.py:332-335): the text
    embeddings, ``bert.encoder.layer.*`` and ``cls.predictions.*``.  (The file also lists ``bert.pooler`` / ``cls.seq_relationship``
    / ``inconsistency_head`` names that no parameter of this model carries, so ``t_pooler`` and ``bi_seq_relationship`` fall to the
    vision learning rate — reproduced here.)"""
    return (name.startswith("bert.embeddings.") or name.startswith("bert.encoder.layer.") or
            (name.startswith("cls.predictions.") and name != "cls.predictions.decoder.weight"))


def param_group(name: str, unused=()) -> int:
    """0 language + decay, 1 language no decay, 2 vision + decay, 3 vision no decay, 4 no gradient (never updated: the reference's
    AdamW skips parameters whose ``grad`` is None — not even weight decay touches them)."""
    if any(m in name for m in UNUSED_MARKERS) or any(name.startswith(u) for u in unused):
        return 4
    nd = any(s in name for s in NO_DECAY)
    return (0 if is_language_weight(name) else 2) + (1 if nd else 0)


def site_seed(base: int, site: str) -> int:
    """32-bit seed of one dropout call site for one forward: the keep-mask of element i is a pure function of (this seed, i)
    (``drop_keep`` in csrc/common.cuh), so the backward regenerates it instead of storing it."""
    return (zlib.crc32(site.encode()) ^ ((int(base) * 2654435761) & 0xFFFFFFFF)) & 0xFFFFFFFF


def warmup_linear_nonzero(step: int, base_lr: float, warmup_steps: int = 10000, t_total: int = 200000, min_lr: float = 1e-5) -> float:
    """utils/optim_utils.py:19-26 (WarmupLinearScheduleNonZero.get_lr) for one base learning rate; ``step`` = scheduler.last_epoch."""
    if step < warmup_steps:
        f = float(step) / float(max(1, warmup_steps))
    else:
        f = max(0.0, float(t_total - step) / float(max(1.0, t_total - warmup_steps)))
    return base_lr * f if base_lr * f > min_lr else min_lr


class ParamStore:
    """Flat fp32 master parameters, gradients, Adam moments and 16-bit operand copies with per-name views."""

    def __init__(self, cfg: ViLBertConfig, ops, unused=()):
        self.cfg, self.ops, self.unused = cfg, ops, tuple(unused)
        shapes = param_shapes(cfg)
        by_group = {g: [] for g in range(5)}
        for name, shp in shapes.items():
            if name in TIED:
                continue
            by_group[param_group(name, unused)].append((name, tuple(shp)))
        self.entries: "OrderedDict[str, tuple]" = OrderedDict()     # name -> (offset, padded shape, real shape)
        self.group_range = {}
        off = 0
        for g in range(5):
            start = off
            for name, shp in by_group[g]:
                pad = tuple(p if p is not None else s for p, s in zip(PADDED[name], shp)) if name in PADDED else shp
                n = int(np.prod(pad))
                self.entries[name] = (off, pad, shp)
                off += (n + 63) // 64 * 64
            self.group_range[g] = (start, off)
        self.total = off
        self.p = ops.zeros32(off)
        self.g = ops.zeros32(off)
        self.m = ops.zeros32(off)
        self.v = ops.zeros32(off)
        self.p16 = ops.empty16(off)
        self.real_shapes = shapes

    def _view(self, flat, name, padded=True):
        name = TIED.get(name, name)
        off, pad, shp = self.entries[name]
        t = flat[off:off + int(np.prod(pad))].view(*pad)
        if not padded and pad != shp:
            t = t[tuple(slice(0, s) for s in shp)]
        return t

    def P(self, name):
        return self._view(self.p, name)

    def G(self, name):
        return self._view(self.g, name)

    def P16(self, name):
        return self._view(self.p16, name)

    def span(self, flat, first, last):
        """One matrix over the adjacent tensors first..last (fused Q | K | V): [sum of rows, K] or [sum of lengths]."""
        o0, p0, _ = self.entries[first]
        o1, p1, _ = self.entries[last]
        n = o1 + int(np.prod(p1)) - o0
        return flat[o0:o0 + n].view(-1, p0[1]) if len(p0) == 2 else flat[o0:o0 + n]

    def load_state_dict(self, sd: Dict[str, torch.Tensor]):
        for name, (off, pad, shp) in self.entries.items():
            src = sd[name].to(torch.float32)
            assert tuple(src.shape) == shp, (name, tuple(src.shape), shp)
            dst = self._view(self.p, name, padded=False)
            dst.copy_(src.to(dst.device, dst.dtype))
        for a, b in TIED.items():
            if a in sd and not torch.equal(sd[a], sd[b]):
                raise ValueError(f"{a} and {b} are tied in the reference but differ in this state dict")
        self.refresh_lp()

    def refresh_lp(self):
        """The 16-bit operand copies of every parameter from the fp32 masters (after a load, or after an optimizer that is not
        ``unimm_t_adamw`` updated the masters in place)."""
        n4 = self.total // 4 * 4
        self.ops.cast_into(self.p[:n4], self.p16[:n4])

    def state_dict(self) -> Dict[str, torch.Tensor]:
        out = OrderedDict()
        for name in self.real_shapes:
            out[name] = self._view(self.p, name, padded=False).detach().clone().cpu().float()
        return out

    def grad_dict(self) -> Dict[str, torch.Tensor]:
        out = OrderedDict()
        for name in self.entries:
            out[name] = self._view(self.g, name, padded=False).detach().clone().cpu()
        return out


class TrainStep:
    """``step(batch)`` = train.py:445-463 for one batch.  ``batch``: CPU tensors ``tokens / segments / positions / labels / weights``
    ``[B, S]`` int64, ``desc [B, 4]`` int32 (``unimm_b200.descriptors``), ``next_sentence_label [B]``, and the image side either per
    sequence (``image_feat [B, R, F]`` ...) or per image plus ``seq_image [B]`` (one block per image, as the loader holds them):
    ``image_feat, image_loc [., R, 5], image_mask [., R], image_label [., R], image_target [., R, C]``; optional ``nsp_weight [2]``."""

    def __init__(self, cfg: ViLBertConfig, state_dict: Dict[str, torch.Tensor], ops, lr: float = 2e-5, image_lr: float = 2e-5,
                 weight_decay: float = 0.01, betas=(0.9, 0.999), eps: float = 1e-6, lm_coeff: float = 1.0, nsp_coeff: float = 1.0,
                 img_coeff: float = 1.0, warmup_steps: int = 10000, t_total: int = 200000, batch_multiply: int = 1, process_group=None,
                 dropout: float = 0.0, seed: int = 0):
        cfg.validate()
        self.cfg, self.ops = cfg, ops
        # a loss whose coefficient is 0 is not part of ``loss`` at all (dense_annotation_finetuning.py:289-293 drops the image term):
        # its head gets no gradient, and parameters without a gradient are skipped by the optimizer
        self.params = ParamStore(cfg, ops, unused=("cls.imagePredictions.",) if img_coeff == 0 else ())
        self.params.load_state_dict(state_dict)
        self.lr, self.image_lr, self.weight_decay, self.betas, self.eps = lr, image_lr, weight_decay, betas, eps
        self.coeff = (lm_coeff, nsp_coeff, img_coeff)
        self.warmup_steps, self.t_total, self.batch_multiply = warmup_steps, t_total, batch_multiply
        # data parallel (true DDP instead of the reference's DataParallelImbalance scatter / replicate / gather,
        # utils/data_parallel.py): one process per GPU, every rank runs the step on ITS sequences, the flat gradient buffer is summed
        # with ONE all-reduce (NCCL over NVLink; gloo in the CPU tests) and the 1 / world factor is folded into AdamW.  The reference
        # normalises every loss per replica and averages the replicas (models/vilbert_dialog.py:1592-1595, train.py:164): the same
        # mean of per-replica gradients.
        self.group = process_group
        self.world = 1
        if process_group is not None or (torch.distributed.is_available() and torch.distributed.is_initialized()):
            self.world = torch.distributed.get_world_size(process_group)
        self.iter_id, self._acc = 0, None
        # every nn.Dropout of the reference's training mode (hidden 0.1, attention probabilities 0.1, the fused pooled vector 0.1:
        # models/vilbert_dialog.py:355, :405, :424, :467, :534, :553, :596, :693, :716, :746, :749, :1065, :1491) with ONE probability;
        # 0 = the reference in eval mode.  Masks are counter-based (site, forward number, element index), regenerated in the backward.
        self.dropout, self.seed, self._forwards, self._seed_base = float(dropout), int(seed), 0, 0
        self.opt_step = 0            # AdamW state['step']
        self.sched_step = 0          # scheduler.last_epoch
        self.S, self.R = None, None

    # ------------------------------------------------------------------ inputs
    def upload(self, batch):
        """Host batch -> device inputs (one H2D per tensor from the caller's (pinned) memory, a few index arrays built on the host)."""
        ops, cfg = self.ops, self.cfg
        dev, fdt = self.params.p.device, self.params.p.dtype       # fdt: fp32 on the device (fp64 only under the CPU test's torch ops)
        t = lambda x, dt=None: torch.as_tensor(np.asarray(x) if not torch.is_tensor(x) else x)          # noqa: E731
        up = lambda x, dt: t(x).to(dt).contiguous().to(dev, non_blocking=True)                          # noqa: E731
        tokens = t(batch["tokens"])
        B, S = tokens.shape
        labels, weights = t(batch["labels"]).long().cpu(), t(batch["weights"]).long().cpu()     # host side: they size the LM-head buffers
        R = t(batch["image_mask"]).shape[-1]
        seq_image = batch.get("seq_image")
        n_img = t(batch["image_feat"]).shape[0]
        if seq_image is None:
            seq_image_np = np.arange(B, dtype=np.int64)
            assert n_img == B
        else:
            seq_image_np = np.asarray(t(seq_image).cpu()).astype(np.int64)
        d = {"B": B, "S": S, "R": R}
        d["ids"], d["seg"], d["pos"] = up(tokens, torch.int64), up(batch["segments"], torch.int64), up(batch["positions"], torch.int64)
        d["desc"] = up(batch["desc"], torch.int32)
        # labelled rows (host side: the loader's tensors are on the host anyway, and the row list sizes every LM-head buffer)
        lab = labels.numpy().reshape(-1)
        rows = np.flatnonzero(lab != -1)
        w = weights.numpy().reshape(-1)
        d["lm_rows"] = up(rows.astype(np.int32), torch.int32)
        d["lm_labels"] = up(lab[rows].astype(np.int32), torch.int32)
        d["lm_weight"] = up(w[rows].astype(np.float32), fdt)
        d["lm_denom"] = float((w != 0).sum())                                    # (lm_weight != 0).sum(), :1594
        d["n_lm"] = int(rows.size)
        d["cls_rows"] = up(np.arange(B, dtype=np.int32) * S, torch.int32)
        d["img0_rows"] = up(np.arange(B, dtype=np.int32) * R, torch.int32)
        row_of = (seq_image_np[:, None] * R + np.arange(R)[None, :]).reshape(-1).astype(np.int32)      # image row of every (sequence, region)
        d["img_row_of"] = up(row_of, torch.int32)
        d["feat"] = up(t(batch["image_feat"]).reshape(n_img * R, -1), fdt)
        loc = t(batch["image_loc"]).float().reshape(n_img * R, -1)
        loc64 = torch.zeros(B * R, 64, device=loc.device)
        loc64[:, :loc.shape[1]] = loc[torch.from_numpy(row_of).long().to(loc.device)]
        d["loc64"] = up(loc64, fdt)
        pick = lambda x: x[torch.from_numpy(seq_image_np).to(x.device)]                                 # noqa: E731
        d["img_mask"] = up(pick(t(batch["image_mask"]).float()), fdt)
        d["img_label"] = up(pick(t(batch["image_label"]).long()).reshape(-1), torch.int64)
        d["img_target"] = up(t(batch["image_target"]).reshape(n_img * R, -1), fdt)
        d["nsl"] = up(batch["next_sentence_label"], torch.int64)
        nw = batch.get("nsp_weight")
        d["nsp_weight"] = None if nw is None else up(t(nw).reshape(-1)[:2], fdt)
        rel = batch.get("gt_relevance")           # dense-annotation fine-tuning: [slates, options] with slates * options == B, in batch order
        if rel is not None:
            rel = t(rel).float()
            assert rel.dim() == 2 and rel.numel() == B, "gt_relevance: [slates, options] covering the batch"
            d["relevance"] = up(rel, fdt)
        return d

    def _drop(self, site: str):
        return None if self.dropout <= 0.0 else (site_seed(self._seed_base, site), self.dropout)

    # ------------------------------------------------------------------ layer pieces (forward saves, backward consumes)
    def _ffn_fwd(self, x32, x16, p_int, p_out, sv):
        """intermediate.dense -> erf-GELU -> output.dense (+ x) -> LayerNorm  (models/vilbert_dialog.py:452-469)."""
        ops, P = self.ops, self.params
        # one epilogue: the pre-activation in fp32 (kept for the backward) and GELU of it as the 16-bit operand of the next GEMM
        t, g16 = ops.linear(x16, P.P16(p_int + ".dense.weight"), P.P(p_int + ".dense.bias"), act=ACT_GELU, want16=True, pre_act32=True)
        d_out = self._drop(p_out)
        pre, _ = ops.linear(g16, P.P16(p_out + ".dense.weight"), P.P(p_out + ".dense.bias"), residual=x32, drop=d_out)
        y32, y16 = ops.layernorm(pre, P.P(p_out + ".LayerNorm.weight"), P.P(p_out + ".LayerNorm.bias"))
        sv.update(ffn_x16=x16, ffn_t=t, ffn_g16=g16, ffn_pre=pre, ffn_drop=d_out)
        return y32, y16

    def _ffn_bwd(self, dy, p_int, p_out, sv):
        ops, P = self.ops, self.params
        d_pre = ops.layernorm_backward(dy, sv["ffn_pre"], P.P(p_out + ".LayerNorm.weight"), P.G(p_out + ".LayerNorm.weight"), P.G(p_out + ".LayerNorm.bias"))
        dg = ops.linear_backward(d_pre, sv["ffn_g16"], P.P16(p_out + ".dense.weight"), P.G(p_out + ".dense.weight"), P.G(p_out + ".dense.bias"),
                                 dx_amax=True, drop=sv["ffn_drop"])
        # GELU' is applied inside the pass that turns dg into the 16-bit operand of the next two GEMMs.  The residual branch's gradient
        # is d_pre itself: the FFN branch is accumulated onto it.
        return ops.linear_backward(dg, sv["ffn_x16"], P.P16(p_int + ".dense.weight"), P.G(p_int + ".dense.weight"), P.G(p_int + ".dense.bias"),
                                   dx_accum=d_pre, gelu_t=sv["ffn_t"])

    def _self_layer_fwd(self, p, x32, x16, B, S, heads, mask_kind, desc, key_mask, sv):
        """BertLayer / BertImageLayer (:479-483, :608-612)."""
        ops, P = self.ops, self.params
        a = p + "attention."
        H = x32.shape[1]
        D = H // heads
        wqkv = P.span(P.p16, a + "self.query.weight", a + "self.value.weight")
        bqkv = P.span(P.p, a + "self.query.bias", a + "self.value.bias")
        _, qkv16 = ops.linear(x16, wqkv, bqkv, want32=False, want16=True)
        d_probs, d_att = self._drop(a + "probs"), self._drop(a + "output")
        ctx16, lse = ops.attention(qkv16[:, :H], qkv16[:, H:2 * H], qkv16[:, 2 * H:], B, heads, D, S, S, mask_kind, desc, key_mask, drop=d_probs)
        pre1, _ = ops.linear(ctx16, P.P16(a + "output.dense.weight"), P.P(a + "output.dense.bias"), residual=x32, drop=d_att)
        y1, y1_16 = ops.layernorm(pre1, P.P(a + "output.LayerNorm.weight"), P.P(a + "output.LayerNorm.bias"))
        sv.update(x16=x16, qkv16=qkv16, ctx16=ctx16, lse=lse, pre1=pre1, dims=(B, S, heads, D, H, mask_kind), drop_probs=d_probs, drop_att=d_att)
        return self._ffn_fwd(y1, y1_16, p + "intermediate", p + "output", sv)

    def _self_layer_bwd(self, p, dy, desc, key_mask, sv):
        ops, P = self.ops, self.params
        a = p + "attention."
        B, S, heads, D, H, mask_kind = sv["dims"]
        dy1 = self._ffn_bwd(dy, p + "intermediate", p + "output", sv)
        d_pre1 = ops.layernorm_backward(dy1, sv["pre1"], P.P(a + "output.LayerNorm.weight"), P.G(a + "output.LayerNorm.weight"),
                                        P.G(a + "output.LayerNorm.bias"))
        dctx = ops.linear_backward(d_pre1, sv["ctx16"], P.P16(a + "output.dense.weight"), P.G(a + "output.dense.weight"), P.G(a + "output.dense.bias"),
                                   dx_amax=True, drop=sv["drop_att"])
        qkv16 = sv["qkv16"]
        dqkv, cell = ops.empty32(qkv16.shape[0], 3 * H), ops.new_amax_cell()
        ops.attention_backward(qkv16[:, :H], qkv16[:, H:2 * H], qkv16[:, 2 * H:], sv["ctx16"], sv["lse"], dctx, B, heads, D, S, S, mask_kind, desc,
                               key_mask, dqkv[:, :H], dqkv[:, H:2 * H], dqkv[:, 2 * H:], cell, drop=sv["drop_probs"])
        ops.register_amax(dqkv, cell)
        wqkv = P.span(P.p16, a + "self.query.weight", a + "self.value.weight")
        return ops.linear_backward(dqkv, sv["x16"], wqkv, P.span(P.g, a + "self.query.weight", a + "self.value.weight"),
                                   P.span(P.g, a + "self.query.bias", a + "self.value.bias"), dx_accum=d_pre1)

    def _conn_layer_fwd(self, p, xv32, xv16, xt32, xt16, inp, sv):
        """BertConnectionLayer (:770-783): stream 1 = image, stream 2 = text; the contexts are swapped by BertBiOutput (:744-754)."""
        ops, P, cfg = self.ops, self.params, self.cfg
        b = p + "biattention."
        Hb, heads = cfg.bi_hidden_size, cfg.bi_num_attention_heads
        D = Hb // heads
        B, S, R = inp["B"], inp["S"], inp["R"]
        _, qkv1 = ops.linear(xv16, P.span(P.p16, b + "query1.weight", b + "value1.weight"), P.span(P.p, b + "query1.bias", b + "value1.bias"),
                             want32=False, want16=True)
        _, qkv2 = ops.linear(xt16, P.span(P.p16, b + "query2.weight", b + "value2.weight"), P.span(P.p, b + "query2.bias", b + "value2.bias"),
                             want32=False, want16=True)
        # text queries over image keys (image padding mask), image queries over text keys (co-attention interval)
        o = p + "biOutput."
        dr = {k: self._drop(n) for k, n in (("p1", b + "probs1"), ("p2", b + "probs2"), ("d1", o + "dense1"), ("d2", o + "dense2"))}
        ctx_t, lse_t = ops.attention(qkv2[:, :Hb], qkv1[:, Hb:2 * Hb], qkv1[:, 2 * Hb:], B, heads, D, S, R, MASK_KEY_VECTOR, None, inp["img_mask"],
                                     drop=dr["p1"])
        ctx_v, lse_v = ops.attention(qkv1[:, :Hb], qkv2[:, Hb:2 * Hb], qkv2[:, 2 * Hb:], B, heads, D, R, S, MASK_CO_INTERVAL, inp["desc"], None,
                                     drop=dr["p2"])
        pre_v, _ = ops.linear(ctx_v, P.P16(o + "dense1.weight"), P.P(o + "dense1.bias"), residual=xv32, drop=dr["d1"])
        av32, av16 = ops.layernorm(pre_v, P.P(o + "LayerNorm1.weight"), P.P(o + "LayerNorm1.bias"))
        pre_t, _ = ops.linear(ctx_t, P.P16(o + "dense2.weight"), P.P(o + "dense2.bias"), residual=xt32, drop=dr["d2"])
        at32, at16 = ops.layernorm(pre_t, P.P(o + "LayerNorm2.weight"), P.P(o + "LayerNorm2.bias"))
        sv.update(xv16=xv16, xt16=xt16, qkv1=qkv1, qkv2=qkv2, ctx_t=ctx_t, lse_t=lse_t, ctx_v=ctx_v, lse_v=lse_v, pre_v=pre_v, pre_t=pre_t,
                  v={}, t={}, dr=dr)
        yv32, yv16 = self._ffn_fwd(av32, av16, p + "v_intermediate", p + "v_output", sv["v"])
        yt32, yt16 = self._ffn_fwd(at32, at16, p + "t_intermediate", p + "t_output", sv["t"])
        return yv32, yv16, yt32, yt16

    def _conn_layer_bwd(self, p, dyv, dyt, inp, sv):
        ops, P, cfg = self.ops, self.params, self.cfg
        b, o = p + "biattention.", p + "biOutput."
        Hb, heads = cfg.bi_hidden_size, cfg.bi_num_attention_heads
        D = Hb // heads
        B, S, R = inp["B"], inp["S"], inp["R"]
        dxv, dxt, dctx_v, dctx_t = None, None, None, None
        if dyv is not None:
            dav = self._ffn_bwd(dyv, p + "v_intermediate", p + "v_output", sv["v"])
            dxv = ops.layernorm_backward(dav, sv["pre_v"], P.P(o + "LayerNorm1.weight"), P.G(o + "LayerNorm1.weight"), P.G(o + "LayerNorm1.bias"))
            dctx_v = ops.linear_backward(dxv, sv["ctx_v"], P.P16(o + "dense1.weight"), P.G(o + "dense1.weight"), P.G(o + "dense1.bias"), dx_amax=True,
                                         drop=sv["dr"]["d1"])
        dat = self._ffn_bwd(dyt, p + "t_intermediate", p + "t_output", sv["t"])
        dxt = ops.layernorm_backward(dat, sv["pre_t"], P.P(o + "LayerNorm2.weight"), P.G(o + "LayerNorm2.weight"), P.G(o + "LayerNorm2.bias"))
        dctx_t = ops.linear_backward(dxt, sv["ctx_t"], P.P16(o + "dense2.weight"), P.G(o + "dense2.weight"), P.G(o + "dense2.bias"), dx_amax=True,
                                     drop=sv["dr"]["d2"])
        qkv1, qkv2 = sv["qkv1"], sv["qkv2"]
        dqkv1, dqkv2 = ops.empty32(qkv1.shape[0], 3 * Hb), ops.empty32(qkv2.shape[0], 3 * Hb)
        cell = ops.new_amax_cell()         # one bound for both matrices: each of the two attentions fills column blocks of both
        ops.attention_backward(qkv2[:, :Hb], qkv1[:, Hb:2 * Hb], qkv1[:, 2 * Hb:], sv["ctx_t"], sv["lse_t"], dctx_t, B, heads, D, S, R,
                               MASK_KEY_VECTOR, None, inp["img_mask"], dqkv2[:, :Hb], dqkv1[:, Hb:2 * Hb], dqkv1[:, 2 * Hb:], cell, drop=sv["dr"]["p1"])
        if dctx_v is None:              # no gradient reaches the image stream's output of this layer (cannot happen with the NSP / image losses on)
            dctx_v = ops.zeros32(qkv1.shape[0], Hb)
            dxv = ops.zeros32(qkv1.shape[0], sv["xv16"].shape[1])
        ops.attention_backward(qkv1[:, :Hb], qkv2[:, Hb:2 * Hb], qkv2[:, 2 * Hb:], sv["ctx_v"], sv["lse_v"], dctx_v, B, heads, D, R, S,
                               MASK_CO_INTERVAL, inp["desc"], None, dqkv1[:, :Hb], dqkv2[:, Hb:2 * Hb], dqkv2[:, 2 * Hb:], cell, drop=sv["dr"]["p2"])
        ops.register_amax(dqkv1, cell)
        ops.register_amax(dqkv2, cell)
        dxv = ops.linear_backward(dqkv1, sv["xv16"], P.span(P.p16, b + "query1.weight", b + "value1.weight"),
                                  P.span(P.g, b + "query1.weight", b + "value1.weight"), P.span(P.g, b + "query1.bias", b + "value1.bias"), dx_accum=dxv)
        dxt = ops.linear_backward(dqkv2, sv["xt16"], P.span(P.p16, b + "query2.weight", b + "value2.weight"),
                                  P.span(P.g, b + "query2.weight", b + "value2.weight"), P.span(P.g, b + "query2.bias", b + "value2.bias"), dx_accum=dxt)
        return dxv, dxt

    # ------------------------------------------------------------------ data parallel: gradient all-reduce overlapped with the backward
    def _layer_ranges(self, prefix: str):
        """Flat-buffer ranges (one per parameter group) of the parameters whose names start with ``prefix``: a layer's tensors are
        consecutive inside each group."""
        P = self.params
        spans = {}
        for name, (off, pad, _) in P.entries.items():
            if name.startswith(prefix):
                g = param_group(name, P.unused)
                if g == 4:
                    continue
                lo, hi = spans.get(g, (off, off))
                spans[g] = (min(lo, off), max(hi, off + (int(np.prod(pad)) + 63) // 64 * 64))
        return list(spans.values())

    def _reduce_async(self, prefix: str, st: dict) -> None:
        """The gradients of the layer ``prefix`` are final: start their all-reduce now (NCCL runs it on its own stream beside the rest of
        the backward); ``backward`` waits for all of them before it returns."""
        if self.world <= 1:
            return
        self.ops.join_side()                     # the layer's wgrad GEMMs run on a second stream (train_ops.linear_backward)
        for a, b in self._layer_ranges(prefix):
            st["reduced"].append((a, b))
            st["handles"].append(torch.distributed.all_reduce(self.params.g[a:b], group=self.group, async_op=True))

    # ------------------------------------------------------------------ the step
    def forward(self, batch=None, inp=None, image_head: Optional[bool] = None) -> dict:
        """The forward of the step: embeddings, encoder (activations saved), the three heads and their loss VALUES, plus the gradients of
        each loss with respect to its head's output at unit coefficient (the fused loss kernels produce them in the same pass).  Returns
        the state ``backward`` consumes; ``state["out"]`` holds the losses as one-element device tensors."""
        ops, P, cfg = self.ops, self.params, self.cfg
        if inp is None:
            inp = self.upload(batch)
        if image_head is None:
            image_head = self.coeff[2] != 0
        ops.begin_step()
        self._seed_base = self.seed + self._forwards            # a fresh set of dropout masks per forward
        self._forwards += 1
        B, S, R = inp["B"], inp["S"], inp["R"]
        P.g.zero_()
        saved = []
        # ---- embeddings (:326-356, :1487-1493)
        e = "bert.embeddings."
        e_sum = ops.embed_text_sum(inp["ids"], inp["seg"], inp["pos"], P.P(e + "word_embeddings.weight"), P.P(e + "position_embeddings.weight"),
                                   P.P(e + "token_type_embeddings.weight"), P.P(e + "token_type_embeddings_extension.weight"), cfg.type_vocab_size)
        xt32, xt16 = ops.layernorm(e_sum, P.P(e + "LayerNorm.weight"), P.P(e + "LayerNorm.bias"))
        d_emb_t, d_emb_v = self._drop("emb.txt"), self._drop("emb.img")
        if d_emb_t is not None:
            xt32, xt16 = ops.dropout(xt32, d_emb_t)
        ve = "bert.v_embeddings."
        feat32 = ops.gather_rows(inp["feat"], inp["img_row_of"])                      # one block per image -> one per sequence
        feat16 = ops.to_lp(feat32)
        del feat32
        loc16 = ops.to_lp(inp["loc64"])
        loc_term = ops.linear_f32(inp["loc64"], P.P(ve + "image_location_embeddings.weight"), P.P(ve + "image_location_embeddings.bias"))   # K padded 5 -> 64
        v_sum, _ = ops.linear(feat16, P.P16(ve + "image_embeddings.weight"), P.P(ve + "image_embeddings.bias"), residual=loc_term)
        xv32, xv16 = ops.layernorm(v_sum, P.P(ve + "LayerNorm.weight"), P.P(ve + "LayerNorm.bias"))
        if d_emb_v is not None:
            xv32, xv16 = ops.dropout(xv32, d_emb_v)
        # ---- encoder (:842-929)
        for kind, i in cfg.layer_schedule():
            sv = {"kind": kind, "i": i}
            if kind == "t":
                xt32, xt16 = self._self_layer_fwd(f"bert.encoder.layer.{i}.", xt32, xt16, B, S, cfg.num_attention_heads, MASK_TEXT_SELF, inp["desc"],
                                                  None, sv)
            elif kind == "v":
                xv32, xv16 = self._self_layer_fwd(f"bert.encoder.v_layer.{i}.", xv32, xv16, B, R, cfg.v_num_attention_heads, MASK_KEY_VECTOR, None,
                                                  inp["img_mask"], sv)
            else:
                xv32, xv16, xt32, xt16 = self._conn_layer_fwd(f"bert.encoder.c_layer.{i}.", xv32, xv16, xt32, xt16, inp, sv)
            saved.append(sv)
        st = {"inp": inp, "saved": saved, "e_sum": e_sum, "v_sum": v_sum, "feat16": feat16, "loc16": loc16, "xv16": xv16, "handles": [],
              "reduced": [], "drop_emb": (d_emb_t, d_emb_v)}
        out = {}
        # ---- masked-LM head + likelihood / unlikelihood loss (:982-986, :1023-1026, :1577-1595), labelled rows only.  The fused vocabulary
        # kernel is forward AND backward of the decoder + loss in one: it runs here with the loss's own normalisation (1 / #weighted tokens)
        if inp["n_lm"] > 0:
            t_ = "cls.predictions.transform."
            x_lm = ops.gather_rows(xt32, inp["lm_rows"])
            x_lm16 = ops.to_lp(x_lm)
            tt, _ = ops.linear(x_lm16, P.P16(t_ + "dense.weight"), P.P(t_ + "dense.bias"))
            g32, _ = ops.gelu(tt, want32=True, want16=False)
            _, h16 = ops.layernorm(g32, P.P(t_ + "LayerNorm.weight"), P.P(t_ + "LayerNorm.bias"), want32=False)
            dH, logp = ops.lm_head_loss_backward(h16, P.P16("bert.embeddings.word_embeddings.weight"), P.P("cls.predictions.bias"), inp["lm_labels"],
                                                 inp["lm_weight"], 1.0 / inp["lm_denom"], P.G("bert.embeddings.word_embeddings.weight"),
                                                 P.G("cls.predictions.bias"))
            out["lm_loss"] = ops.lm_ul_value(logp, inp["lm_weight"], 1.0 / inp["lm_denom"])
            st.update(lm=(x_lm16, tt, g32, dH))
        else:
            out["lm_loss"] = ops.zeros32(1)
        # ---- poolers + NSP head + weighted CE (:946-967, :1062-1070, :1605-1621)
        cls_t, cls_v = ops.gather_rows(xt32, inp["cls_rows"]), ops.gather_rows(xv32, inp["img0_rows"])
        pt = ops.linear_f32(cls_t, P.P("bert.t_pooler.dense.weight"), P.P("bert.t_pooler.dense.bias"), act=ACT_RELU)
        pv = ops.linear_f32(cls_v, P.P("bert.v_pooler.dense.weight"), P.P("bert.v_pooler.dense.bias"), act=ACT_RELU)
        fused = ops.mul(pt, pv)
        d_pool = self._drop("nsp.pooled")
        if d_pool is not None:
            fused, _ = ops.dropout(fused, d_pool, want16=False)
        nsp_logits = ops.linear_f32(fused, P._view(P.p, "cls.bi_seq_relationship.weight", padded=False),
                                    P._view(P.p, "cls.bi_seq_relationship.bias", padded=False))
        out["nsp_loss"], d_nsp = ops.nsp_ce(nsp_logits, inp["nsl"], inp["nsp_weight"], 1.0)
        st.update(nsp=(cls_t, cls_v, pt, pv, fused, nsp_logits, d_nsp), drop_pool=d_pool)
        # ---- image head + masked KL (:1085-1088, :1569-1574)
        if image_head:
            ih = "cls.imagePredictions."
            tv, _ = ops.linear(xv16, P.P16(ih + "transform.dense.weight"), P.P(ih + "transform.dense.bias"))
            gv32, _ = ops.gelu(tv, want32=True, want16=False)
            _, hv16 = ops.layernorm(gv32, P.P(ih + "transform.LayerNorm.weight"), P.P(ih + "transform.LayerNorm.bias"), want32=False)
            v_logits, _ = ops.linear(hv16, P.P16(ih + "decoder.weight"), P.P(ih + "decoder.bias"))
            out["img_loss"], d_vlog = ops.image_kl(v_logits, cfg.v_target_size, inp["img_target"], inp["img_row_of"], inp["img_label"], 1.0)
            del v_logits
            st.update(img=(tv, gv32, hv16, d_vlog))
        st["out"] = out
        return st

    def backward(self, st: dict, lm_c: float = 1.0, nsp_c: float = 1.0, img_c: float = 1.0, d_nsp_logits=None, ndcg_scale: float = 0.0) -> dict:
        """The backward of ``forward``'s state for ``lm_c lm + nsp_c nsp + img_c img`` (+ ``<d_nsp_logits, nsp logits>`` for a caller that
        differentiates through the NSP scores itself, + ``ndcg_scale neuralNDCG_transposed`` when the batch carries ``gt_relevance``): every
        parameter gradient lands in ``params.g``.  Consumes the state."""
        ops, P, cfg = self.ops, self.params, self.cfg
        inp, saved = st["inp"], st["saved"]
        B, S, R = inp["B"], inp["S"], inp["R"]
        e, ve = "bert.embeddings.", "bert.v_embeddings."
        extra = {}
        d_xt, d_xv = ops.zeros32(B * S, cfg.hidden_size), ops.zeros32(B * R, cfg.v_hidden_size)
        if "lm" in st:
            t_ = "cls.predictions.transform."
            x_lm16, tt, g32, dH = st.pop("lm")
            if lm_c != 1.0:                       # the decoder's gradients were written at unit coefficient by the forward's fused kernel
                ops.ew(EW_SCALE, dH, alpha=lm_c)
                ops.ew(EW_SCALE, P.G("bert.embeddings.word_embeddings.weight").view(-1), alpha=lm_c)
                ops.ew(EW_SCALE, P.G("cls.predictions.bias"), alpha=lm_c)
            dg = ops.layernorm_backward(dH, g32, P.P(t_ + "LayerNorm.weight"), P.G(t_ + "LayerNorm.weight"), P.G(t_ + "LayerNorm.bias"))
            dtt = ops.gelu_backward(dg, tt)
            dx_lm = ops.linear_backward(dtt, x_lm16, P.P16(t_ + "dense.weight"), P.G(t_ + "dense.weight"), P.G(t_ + "dense.bias"))
            ops.scatter_add_rows(dx_lm, inp["lm_rows"], d_xt)
        cls_t, cls_v, pt, pv, fused, nsp_logits, d_nsp = st.pop("nsp")
        if nsp_c != 1.0:
            ops.ew(EW_SCALE, d_nsp.view(-1), alpha=nsp_c)
        if d_nsp_logits is not None:
            ops.ew(EW_ADD, d_nsp.view(-1), d_nsp_logits.reshape(-1).contiguous())
        if "relevance" in inp and ndcg_scale != 0.0:
            # dense-annotation objective (dense_annotation_finetuning.py:267-293): neuralNDCG_transposed on y_pred = softmax(nsp)[:, 0]
            p0 = ops.nsp_prob0(nsp_logits)
            d_p0, extra["ndcg"] = ops.neural_ndcg_backward(p0.view(*inp["relevance"].shape), inp["relevance"], ndcg_scale)
            ops.nsp_prob0_backward(nsp_logits, d_p0.view(-1), d_nsp)
        d_nsp64 = ops.zeros32(B, 64)
        d_nsp64[:, :2].copy_(d_nsp)
        dfused = ops.linear_backward(d_nsp64, ops.to_lp(fused), P.P16("cls.bi_seq_relationship.weight"), P.G("cls.bi_seq_relationship.weight"),
                                     P.G("cls.bi_seq_relationship.bias"))
        if st["drop_pool"] is not None:
            ops.dropout_backward(dfused, st["drop_pool"])
        dpt, dpv = ops.relu_backward(ops.mul(dfused, pv), pt), ops.relu_backward(ops.mul(dfused, pt), pv)
        dcls_t = ops.linear_backward(dpt, ops.to_lp(cls_t), P.P16("bert.t_pooler.dense.weight"), P.G("bert.t_pooler.dense.weight"),
                                     P.G("bert.t_pooler.dense.bias"))
        dcls_v = ops.linear_backward(dpv, ops.to_lp(cls_v), P.P16("bert.v_pooler.dense.weight"), P.G("bert.v_pooler.dense.weight"),
                                     P.G("bert.v_pooler.dense.bias"))
        ops.scatter_add_rows(dcls_t, inp["cls_rows"], d_xt)
        ops.scatter_add_rows(dcls_v, inp["img0_rows"], d_xv)
        if "img" in st and img_c != 0.0:
            ih = "cls.imagePredictions."
            tv, gv32, hv16, d_vlog = st.pop("img")
            if img_c != 1.0:
                ops.ew(EW_SCALE, d_vlog.view(-1), alpha=img_c)
            dhv = ops.linear_backward(d_vlog, hv16, P.P16(ih + "decoder.weight"), P.G(ih + "decoder.weight"), P.G(ih + "decoder.bias"))
            dgv = ops.layernorm_backward(dhv, gv32, P.P(ih + "transform.LayerNorm.weight"), P.G(ih + "transform.LayerNorm.weight"),
                                         P.G(ih + "transform.LayerNorm.bias"))
            dtv = ops.gelu_backward(dgv, tv)
            d_xv = ops.linear_backward(dtv, st["xv16"], P.P16(ih + "transform.dense.weight"), P.G(ih + "transform.dense.weight"),
                                       P.G(ih + "transform.dense.bias"), dx_accum=d_xv)
            del tv, gv32, hv16, d_vlog, dhv
        # ---- encoder, in reverse
        for sv in reversed(saved):
            kind, i = sv["kind"], sv["i"]
            prefix = {"t": "bert.encoder.layer.", "v": "bert.encoder.v_layer.", "c": "bert.encoder.c_layer."}[kind] + f"{i}."
            if kind == "t":
                d_xt = self._self_layer_bwd(prefix, d_xt, inp["desc"], None, sv)
            elif kind == "v":
                d_xv = self._self_layer_bwd(prefix, d_xv, None, inp["img_mask"], sv)
            else:
                d_xv, d_xt = self._conn_layer_bwd(prefix, d_xv, d_xt, inp, sv)
            sv.clear()
            self._reduce_async(prefix, st)
        # ---- embeddings
        if st["drop_emb"][0] is not None:
            ops.dropout_backward(d_xt, st["drop_emb"][0])
            ops.dropout_backward(d_xv, st["drop_emb"][1])
        d_esum = ops.layernorm_backward(d_xt, st["e_sum"], P.P(e + "LayerNorm.weight"), P.G(e + "LayerNorm.weight"), P.G(e + "LayerNorm.bias"))
        ops.embed_text_backward(d_esum, inp["ids"], inp["seg"], inp["pos"], P.G(e + "word_embeddings.weight"), P.G(e + "position_embeddings.weight"),
                                P.G(e + "token_type_embeddings.weight"), P.G(e + "token_type_embeddings_extension.weight"), cfg.type_vocab_size)
        d_vsum = ops.layernorm_backward(d_xv, st["v_sum"], P.P(ve + "LayerNorm.weight"), P.G(ve + "LayerNorm.weight"), P.G(ve + "LayerNorm.bias"))
        ops.linear_backward(d_vsum, st["feat16"], P.P16(ve + "image_embeddings.weight"), P.G(ve + "image_embeddings.weight"),
                            P.G(ve + "image_embeddings.bias"), need_dx=False)
        ops.linear_backward(d_vsum, st["loc16"], P.P16(ve + "image_location_embeddings.weight"), P.G(ve + "image_location_embeddings.weight"),
                            P.G(ve + "image_location_embeddings.bias"), need_dx=False)
        ops.join_side()
        if self.world > 1:
            # what the per-layer all-reduces above have not covered (embeddings, poolers, heads): the gaps of the range that holds every
            # parameter with a gradient.  Sums; optimizer_step divides by the world size.
            lo, hi = P.group_range[0][0], P.group_range[3][1]
            pos = lo
            for a, b in sorted(st["reduced"]) + [(hi, hi)]:
                if a > pos:
                    st["handles"].append(torch.distributed.all_reduce(P.g[pos:a], group=self.group, async_op=True))
                pos = max(pos, b)
            for h in st["handles"]:
                h.wait()
        st.clear()
        return extra

    def forward_backward(self, batch=None, inp=None, read_losses: bool = True) -> Dict[str, float]:
        """Forward, the three losses, backward: the gradients of ``lm_coeff lm + nsp_coeff nsp + img_coeff img`` (each ``/ batch_multiply``;
        ``+ neuralNDCG_transposed`` when the batch carries ``gt_relevance``) land in ``params.g``.  ``inp``: inputs already on the device
        (``upload``); ``read_losses=False`` returns device tensors instead of floats."""
        st = self.forward(batch, inp)
        out = st["out"]
        lm_c, nsp_c, img_c = (c / self.batch_multiply for c in self.coeff)
        out.update(self.backward(st, lm_c, nsp_c, img_c, ndcg_scale=1.0 / self.batch_multiply))
        if not read_losses:
            return out
        ndcg = out.pop("ndcg", None)
        if self.world > 1:                                                # the mean of the replicas' loss values (train.py:164)
            keys = sorted(out)
            packed = torch.cat([out[k].reshape(1) for k in keys])
            torch.distributed.all_reduce(packed, group=self.group)
            out = {k: packed[i:i + 1] / self.world for i, k in enumerate(keys)}
        vals = {k: float(v.item()) for k, v in out.items()}              # the step's device -> host read
        vals.setdefault("lm_loss", 0.0)
        vals.setdefault("img_loss", 0.0)
        vals["loss"] = self.coeff[0] * vals["lm_loss"] + self.coeff[1] * vals["nsp_loss"] + self.coeff[2] * vals["img_loss"]
        if ndcg is not None:                                              # -mean of the per-slate NDCG over the slates with a relevant option
            nd = ndcg.detach().cpu().double()
            nz = nd != 0
            vals["ndcg_loss"] = float(-(nd[nz].sum() / nz.sum())) if bool(nz.any()) else 0.0
            vals["loss"] += vals["ndcg_loss"]
        return vals

    def optimizer_step(self):
        """scaler.step(optimizer); scheduler.step()  (train.py:455-463) with the learning rates of the CURRENT scheduler state."""
        ops, P = self.ops, self.params
        self.opt_step += 1
        lr_l = warmup_linear_nonzero(self.sched_step, self.lr, self.warmup_steps, self.t_total)
        lr_v = warmup_linear_nonzero(self.sched_step, self.image_lr, self.warmup_steps, self.t_total)
        for g, (lr, wd) in enumerate(((lr_l, self.weight_decay), (lr_l, 0.0), (lr_v, self.weight_decay), (lr_v, 0.0))):
            a, b = P.group_range[g]
            if b > a:
                ops.adamw(P.p[a:b], P.g[a:b], P.m[a:b], P.v[a:b], lr, self.betas[0], self.betas[1], self.eps, wd, self.opt_step, True,
                          1.0 / self.world, P.p16[a:b])
        self.sched_step += 1

    def step(self, batch=None, inp=None, read_losses: bool = True) -> Dict[str, float]:
        """One iteration of train.py:445-463.  With ``batch_multiply`` k > 1 the gradients of k calls are summed (every loss already carries
        1 / k) and the optimizer steps when ``iter_id % k == 0`` — iteration 0 included, as the reference's ``or iter_id == 0`` has it —
        while the scheduler advances on every call (train.py:455-463)."""
        vals = self.forward_backward(batch, inp, read_losses)
        k = self.batch_multiply
        if k > 1:
            P, ops = self.params, self.ops
            if self._acc is None:
                self._acc = ops.zeros32(P.total)
            ops.ew(EW_ADD, self._acc, P.g)
            if self.iter_id % k == 0:
                P.g.copy_(self._acc)
                self._acc.zero_()
                self.optimizer_step()
            else:
                self.sched_step += 1
        else:
            self.optimizer_step()
        self.iter_id += 1
        return vals

    # ------------------------------------------------------------------ checkpoint / resume (train.py:503-505 saves model, optimizer, scheduler)
    def optimizer_state_dict(self) -> dict:
        """Adam moments per parameter (reference key layout, real shapes, fp32 on the host) and the step counters."""
        P = self.params
        return {"exp_avg": {n: P._view(P.m, n, padded=False).detach().clone().cpu() for n in P.entries},
                "exp_avg_sq": {n: P._view(P.v, n, padded=False).detach().clone().cpu() for n in P.entries},
                "step": self.opt_step, "scheduler_last_epoch": self.sched_step, "iter_id": self.iter_id}

    def load_optimizer_state_dict(self, sd: dict) -> None:
        P = self.params
        for n in P.entries:
            for flat, key in ((P.m, "exp_avg"), (P.v, "exp_avg_sq")):
                dst = P._view(flat, n, padded=False)
                dst.copy_(sd[key][n].to(dst.device, dst.dtype))
        self.opt_step, self.sched_step, self.iter_id = int(sd["step"]), int(sd["scheduler_last_epoch"]), int(sd.get("iter_id", 0))
        if self._acc is not None:
            self._acc.zero_()

    def state_dict(self):
        return self.params.state_dict()

    def grad_dict(self):
        return self.params.grad_dict()
