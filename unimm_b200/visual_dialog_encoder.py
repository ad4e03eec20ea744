"""Drop-in for the reference's ``VisualDialogEncoder`` (reference models/visual_dialog_encoder.py:8-50).

Same constructor argument (the model JSON), same ``forward`` keyword set and return tuple, same 535-key
``state_dict`` under the ``bert_pretrained.`` prefix — so ``train.forward`` (reference train.py:142-161)
and a reference checkpoint work unchanged — but the body runs in ``libunimm_b200.so``.

Differences that are deliberate and visible:
  * by default inference only (forward + losses) — after ``enable_training()`` the loss branch IS differentiable: the three losses (and the
    NSP scores) come out of a ``torch.autograd.Function`` whose backward runs the device backward of ``unimm_b200.train_step`` and hands
    every parameter its gradient, so the reference's own loop (``scaler.scale(loss).backward(); scaler.step(optimizer)``,
    train.py:453-463, dense_annotation_finetuning.py:253-300) trains through these kernels with ANY torch optimizer;
  * dropout is the identity on the inference paths (the reference's ``.eval()`` behaviour); the differentiable branch applies it in
    ``train()`` mode with this library's counter-based masks (not torch's RNG stream);
  * dense masks are converted to 4-integer descriptors and verified (sequences truncated at max_seq_len included); any other
    mask pattern raises;
  * ``score()`` is the fast entry: per-sequence log-likelihoods without the [B,S,30522] logits that
    ``output_lm_scores=True`` has to materialise for compatibility.
"""
from __future__ import annotations

from typing import Optional

import torch
from torch import nn

from .config import ViLBertConfig
from .descriptors import descriptors_from_masks
from .engine import Engine
from .weights import TIED, param_shapes


def _build_param_tree(root: nn.Module, cfg: ViLBertConfig) -> None:
    """Register parameters under exactly the reference's dotted names (order included)."""
    made = {}
    for name, shape in param_shapes(cfg).items():
        *path, leaf = name.split(".")
        mod = root
        for p in path:
            if p not in mod._modules:
                mod.add_module(p, nn.Module())
            mod = mod._modules[p]
        if name in TIED:
            param = made[TIED[name]]          # decoder.weight is the word-embedding Parameter (ref :1020)
        else:
            param = nn.Parameter(torch.zeros(shape), requires_grad=False)
        made[name] = param
        mod.register_parameter(leaf, param)


class _TrainBridge(torch.autograd.Function):
    """Splices the device training step into torch.autograd: forward = ``TrainStep.forward`` (already run by the module; its state rides on
    the context), backward = ``TrainStep.backward`` with the incoming gradients of the three losses as coefficients (plus the gradient with
    respect to the NSP scores, which dense_annotation_finetuning.py differentiates through itself), returning each parameter's gradient."""

    @staticmethod
    def forward(ctx, module, state, *params):
        ctx.module, ctx.state = module, state
        out, dev = state["out"], state["nsp"][5].device
        img = out["img_loss"] if "img_loss" in out else torch.zeros(1, device=dev)
        return out["lm_loss"].clone(), img.clone(), out["nsp_loss"].clone(), state["nsp"][5].clone()

    @staticmethod
    def backward(ctx, g_lm, g_img, g_nsp, g_scores):
        module, st = ctx.module, ctx.state
        if st is None or "inp" not in st:
            raise RuntimeError("the device training step keeps ONE forward's activations: backward was already run for this forward")
        ctx.state = None
        ts = module._train
        val = lambda g: 0.0 if g is None else float(g.reshape(-1)[0].item())                          # noqa: E731
        lm_c, img_c, nsp_c = val(g_lm), val(g_img), val(g_nsp)
        had_image_head = "img" in st
        ts.backward(st, lm_c, nsp_c, img_c, d_nsp_logits=None if g_scores is None else g_scores.to(torch.float32))
        grads = []
        for name, _ in module._train_params:
            dead = module._train_group[name] == 4 or (name.startswith("cls.imagePredictions.") and (img_c == 0.0 or not had_image_head))
            grads.append(None if dead else ts.params._view(ts.params.g, name, padded=False).clone())
        return (None, None) + tuple(grads)


class VisualDialogEncoder(nn.Module):
    def __init__(self, config_path, precision: str = "fp32", max_sequences: int = 128, device: Optional[int] = None,
                 verify_masks: bool = True):
        super().__init__()
        self.config = config_path if isinstance(config_path, ViLBertConfig) else ViLBertConfig.from_json_file(config_path)
        self.bert_pretrained = nn.Module()
        _build_param_tree(self.bert_pretrained, self.config)
        self.precision, self.max_sequences, self.device_index = precision, max_sequences, device
        self.verify_masks = verify_masks
        self._engine: Optional[Engine] = None
        self._dirty = True
        self._train = None

    # ------------------------------------------------------------------ training (SURVEY.md 8f item 1 behind the reference's own loop)
    def enable_training(self, precision: str = "fp16", device: Optional[int] = None, dropout: float = 0.1):
        """Make the loss branch differentiable.  ``dropout`` (the reference's 0.1 at every nn.Dropout site) is applied while the module is
        in ``train()`` mode and not in ``eval()`` mode, as nn.Dropout behaves.  The module's Parameters become fp32 views of the device training step's flat master
        buffer (same names, shapes and values; ``requires_grad=True``), so an optimizer built from ``named_parameters()`` AFTER this call
        updates the masters in place; every training forward refreshes the 16-bit operand copies from them.  Returns ``self``."""
        from .train_ops import DeviceOps
        from .train_step import TrainStep, param_group
        from .weights import strip_prefix
        dev = torch.device("cuda", torch.cuda.current_device() if device is None else device)
        sd = strip_prefix({k: v.detach().cpu() for k, v in self.state_dict().items()})
        self._train = TrainStep(self.config, sd, DeviceOps(dev, precision))
        self._train_dropout = float(dropout)
        P = self._train.params
        made, self._train_params, self._train_group = {}, [], {}
        for name in param_shapes(self.config):
            *path, leaf = name.split(".")
            mod = self.bert_pretrained
            for p in path:
                mod = mod._modules[p]
            if name in TIED:
                param = made[TIED[name]]
            else:
                param = nn.Parameter(P._view(P.p, name, padded=False), requires_grad=True)
                self._train_params.append((name, param))
                self._train_group[name] = param_group(name)
            made[name] = param
            mod._parameters[leaf] = param
        self._dirty = True
        return self

    def _training_forward(self, input_ids, image_feat, image_loc, token_type_ids, token_position_ids, desc, masked_lm_labels,
                          next_sentence_label, image_attention_mask, image_label, image_target, nsp_weight, lm_weight, output_nsp_scores):
        ts = self._train
        ts.dropout = self._train_dropout if self.training else 0.0
        ts.params.refresh_lp()                                   # the optimizer wrote the fp32 masters since the last forward
        if lm_weight is None:
            raise ValueError("the device training step implements the likelihood / unlikelihood loss: pass lm_weight (train.forward does)")
        batch = {"tokens": input_ids, "segments": token_type_ids, "positions": token_position_ids, "labels": masked_lm_labels,
                 "weights": lm_weight, "desc": desc, "next_sentence_label": next_sentence_label, "image_feat": image_feat,
                 "image_loc": image_loc, "image_mask": image_attention_mask, "image_label": image_label, "image_target": image_target,
                 "nsp_weight": nsp_weight}
        state = ts.forward(batch, image_head=True)
        lm, img, nsp, scores = _TrainBridge.apply(self, state, *[p for _, p in self._train_params])
        self._dirty = True                                       # the inference engine's copy of the weights is stale from here on
        return (lm, img, nsp) + ((scores,) if output_nsp_scores else ())

    # ------------------------------------------------------------------ weights
    def load_state_dict(self, state_dict, strict: bool = True, **kw):
        r = super().load_state_dict(state_dict, strict=strict, **kw)
        self._dirty = True
        return r

    def engine(self) -> Engine:
        if self._engine is None or self._dirty:
            if self._engine is not None:
                self._engine.close()
            self._engine = Engine(self.config, self.state_dict(), precision=self.precision,
                                  max_sequences=self.max_sequences, device=self.device_index)
            self._dirty = False
        return self._engine

    def _masks_on_device(self, attention_mask, co_attention_mask):
        """The dense masks are only read to derive (and verify) the 4-integer descriptors: do that on the engine's device instead of
        with CPU reductions over [B,S,S] (the reference's callers hand over CPU tensors, train.py:113-129).  int64 masks are narrowed
        to bool on the host first (8x fewer bytes over PCIe)."""
        dev = self.engine().device
        if attention_mask.device != dev:
            if attention_mask.dtype not in (torch.bool, torch.uint8):
                attention_mask = attention_mask != 0
            attention_mask = attention_mask.to(dev, non_blocking=True)
        if co_attention_mask.device != dev:
            co_attention_mask = co_attention_mask.to(dev, non_blocking=True)
        return attention_mask, co_attention_mask

    # ------------------------------------------------------------------ fast entry
    @torch.no_grad()
    def score(self, input_ids, image_feat, image_loc, token_type_ids, token_position_ids, masked_lm_labels,
              image_attention_mask, desc=None, attention_mask=None, co_attention_mask=None, feat_index=None,
              want=("seq_score", "nsp_scores")):
        """Per-sequence generative scores (val_lm.py:131-136) and NSP logits; chunks internally."""
        eng = self.engine()
        if desc is None:
            desc = descriptors_from_masks(*self._masks_on_device(attention_mask, co_attention_mask), verify=self.verify_masks)
        B = input_ids.shape[0]
        outs = {k: [] for k in want}
        for s in range(0, B, eng.max_sequences):
            e = min(B, s + eng.max_sequences)
            sl = slice(s, e)
            if feat_index is None:
                f, l, m, fi = image_feat[sl], image_loc[sl], image_attention_mask[sl], None
            else:
                f, l, m, fi = image_feat, image_loc, image_attention_mask, feat_index[sl]
            o = eng.forward(input_ids[sl], token_type_ids[sl], token_position_ids[sl], desc[sl], f, l, m, feat_index=fi,
                            masked_lm_labels=masked_lm_labels[sl], want=want)
            for k in want:
                outs[k].append(o[k])
        eng.check_ids()
        return {k: torch.cat(v, 0) for k, v in outs.items()}

    # ------------------------------------------------------------------ reference signature
    def forward(self, input_ids, image_feat, image_loc, sep_indices=None, sep_len=None, token_type_ids=None,
                token_position_ids=None, attention_mask=None, masked_lm_labels=None, next_sentence_label=None,
                head_mask=None, random_round_indices=None, output_nsp_scores=False, output_lm_scores=False,
                image_attention_mask=None, co_attention_mask=None, image_label=None, image_target=None, nsp_weight=None,
                lm_weight=None):
        differentiable = (self._train is not None and torch.is_grad_enabled() and next_sentence_label is not None and
                          masked_lm_labels is not None and image_target is not None)
        if not differentiable:
            with torch.no_grad():
                return self._inference_forward(input_ids, image_feat, image_loc, token_type_ids, token_position_ids, attention_mask,
                                               masked_lm_labels, next_sentence_label, output_nsp_scores, output_lm_scores,
                                               image_attention_mask, co_attention_mask, image_label, image_target, nsp_weight, lm_weight)
        if output_lm_scores:
            raise NotImplementedError("the training branch does not materialise [B, S, 30522] logits (train.py's training call does not ask)")
        B, S = input_ids.shape
        if token_type_ids is None:
            token_type_ids = torch.zeros_like(input_ids)
        if token_position_ids is None:
            token_position_ids = torch.arange(S, dtype=torch.long, device=input_ids.device).unsqueeze(0).expand(B, S)
        if image_attention_mask is None:
            image_attention_mask = torch.ones(image_feat.shape[:2])
        if attention_mask is None or co_attention_mask is None:
            raise ValueError("attention_mask and co_attention_mask are required (the reference callers always pass them)")
        dev = self._train.params.p.device
        am = attention_mask if attention_mask.dtype in (torch.bool, torch.uint8) else attention_mask != 0
        desc = descriptors_from_masks(am.to(dev, non_blocking=True), co_attention_mask.to(dev, non_blocking=True), verify=self.verify_masks)
        return self._training_forward(input_ids, image_feat, image_loc, token_type_ids, token_position_ids, desc, masked_lm_labels,
                                      next_sentence_label, image_attention_mask, image_label, image_target, nsp_weight, lm_weight,
                                      output_nsp_scores)

    def _inference_forward(self, input_ids, image_feat, image_loc, token_type_ids, token_position_ids, attention_mask, masked_lm_labels,
                           next_sentence_label, output_nsp_scores, output_lm_scores, image_attention_mask, co_attention_mask, image_label,
                           image_target, nsp_weight, lm_weight):
        eng = self.engine()
        B, S = input_ids.shape
        if token_type_ids is None:
            token_type_ids = torch.zeros_like(input_ids)
        if token_position_ids is None:
            token_position_ids = torch.arange(S, dtype=torch.long, device=input_ids.device).unsqueeze(0).expand(B, S)
        if image_attention_mask is None:
            image_attention_mask = torch.ones(image_feat.shape[:2])
        if attention_mask is None or co_attention_mask is None:
            raise ValueError("attention_mask and co_attention_mask are required (the reference callers always pass them)")
        desc = descriptors_from_masks(*self._masks_on_device(attention_mask, co_attention_mask), verify=self.verify_masks)
        training = next_sentence_label is not None and masked_lm_labels is not None and image_target is not None
        want = []
        if output_nsp_scores:
            want.append("nsp_scores")
        if output_lm_scores:
            want.append("prediction_scores_t")
        if training:
            if B > eng.max_sequences:
                raise ValueError(f"the loss branch needs the whole batch in one chunk: construct with max_sequences >= {B}")
            want.append("losses")
            o = eng.forward(input_ids, token_type_ids, token_position_ids, desc, image_feat, image_loc, image_attention_mask,
                            masked_lm_labels=masked_lm_labels, lm_weight=lm_weight, next_sentence_label=next_sentence_label,
                            image_label=image_label, image_target=image_target, nsp_weight=nsp_weight, want=tuple(want))
            losses = o["losses"]
            out = (losses[0:1].clone(), losses[1:2].clone(), losses[2:3].clone())
            chunks = [o]
        else:
            out = (None, None, None)
            chunks = []
            for s in range(0, B, eng.max_sequences):
                sl = slice(s, min(B, s + eng.max_sequences))
                chunks.append(eng.forward(input_ids[sl], token_type_ids[sl], token_position_ids[sl], desc[sl], image_feat[sl],
                                          image_loc[sl], image_attention_mask[sl],
                                          masked_lm_labels=None if masked_lm_labels is None else masked_lm_labels[sl],
                                          want=tuple(want)))
        eng.check_ids()          # an out-of-range id is an error here exactly as in the reference's nn.Embedding
        if output_nsp_scores:
            out = out + (torch.cat([c["nsp_scores"] for c in chunks], 0),)
        if output_lm_scores:
            out = out + (torch.cat([c["prediction_scores_t"] for c in chunks], 0),)
        return out
