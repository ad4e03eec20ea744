"""ORACLE (test infrastructure, not product code): the optimizer and schedule of the reference's training step.

``train.py:22`` imports ``AdamW`` from ``pytorch_transformers.optimization`` — a third-party dependency that is NOT vendored in
/root/reference and not installed in this image (the reference pins no version; the package's last release is 1.2.0).  Its published
algorithm (``AdamW.step``; defaults betas (0.9, 0.999), eps 1e-6, weight_decay 0.0, correct_bias True) is restated below.  Note the
two differences from ``torch.optim.AdamW``: eps is added to the UNcorrected sqrt(v), and the decoupled decay is applied AFTER the
Adam update, on the updated parameter.  Parameters whose ``grad`` is None are skipped entirely.  "parity unpinned" for this file:
there is no reference-side golden vector for the optimizer (the dependency cannot be executed here); the restatement is anchored
on the call site train.py:322-347 (groups: lr by membership in config/language_weights.json, weight_decay 0 where the name contains
'bias' / 'LayerNorm.bias' / 'LayerNorm.weight', else 0.01) and the schedule class utils/optim_utils.py:8-26, which IS in the
reference and is restated by ``warmup_linear_nonzero_lr``.
"""
from __future__ import annotations

import math

import torch


def adamw_step(p: torch.Tensor, grad: torch.Tensor, state: dict, lr: float, betas=(0.9, 0.999), eps: float = 1e-6,
               weight_decay: float = 0.0, correct_bias: bool = True) -> None:
    """One ``AdamW.step`` for one parameter, in place (pytorch_transformers/optimization.py, class AdamW)."""
    if not state:
        state["step"], state["exp_avg"], state["exp_avg_sq"] = 0, torch.zeros_like(p), torch.zeros_like(p)
    state["step"] += 1
    b1, b2 = betas
    state["exp_avg"].mul_(b1).add_(grad, alpha=1.0 - b1)
    state["exp_avg_sq"].mul_(b2).addcmul_(grad, grad, value=1.0 - b2)
    denom = state["exp_avg_sq"].sqrt().add_(eps)
    step_size = lr
    if correct_bias:
        step_size = lr * math.sqrt(1.0 - b2 ** state["step"]) / (1.0 - b1 ** state["step"])
    p.addcdiv_(state["exp_avg"], denom, value=-step_size)
    if weight_decay > 0.0:
        p.add_(p, alpha=-lr * weight_decay)


def warmup_linear_nonzero_lr(last_epoch: int, base_lr: float, warmup_steps: int = 10000, t_total: int = 200000, min_lr: float = 1e-5) -> float:
    """WarmupLinearScheduleNonZero.get_lr for one group (utils/optim_utils.py:19-26)."""
    if last_epoch < warmup_steps:
        f = float(last_epoch) / float(max(1, warmup_steps))
    else:
        f = max(0, float(t_total - last_epoch) / float(max(1.0, t_total - warmup_steps)))
    return base_lr * f if (base_lr * f) > min_lr else min_lr


NO_DECAY = ["bias", "LayerNorm.bias", "LayerNorm.weight"]                     # train.py:323


def group_of(name: str, language_weights, lr: float, image_lr: float):
    """(lr, weight_decay) of one named parameter as train.py:329-345 assigns them; ``name`` carries the model's own prefix
    (``bert_pretrained.``), as the entries of config/language_weights.json do."""
    return (lr if name in language_weights else image_lr), (0.0 if any(nd in name for nd in NO_DECAY) else 0.01)
