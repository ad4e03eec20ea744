"""Dense-annotation objective pieces on the GPU (reference utils/rank_loss.py:518-581, val.py:152-161): forward values and the gradient
of neuralNDCG_transposed with respect to the predicted scores (``neural_ndcg_loss_backward``).

``neural_ndcg_loss(y_pred, y_true)`` is ``neuralNDCG_transposed(y_pred, y_true)`` with the defaults
dense_annotation_finetuning.py:288 uses; ``ensemble_normalise(probs)`` is the 5-model NSP ensemble of val.py / evaluate.py.
"""
from __future__ import annotations

import ctypes as C

import torch

from ._lib import check, lib, ptr


def _stream(dev):
    return C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def neural_ndcg_loss(y_pred: torch.Tensor, y_true: torch.Tensor, temperature: float = 1.0, max_iter: int = 50, tol: float = 1e-6,
                     return_parts: bool = False):
    if not y_pred.is_cuda:
        raise ValueError("neural_ndcg_loss runs on the device the scores live on (CUDA)")
    if (y_true == -1).any():
        raise NotImplementedError("padded slates (y_true == -1) are not part of the dense-annotation path")
    n = y_pred.shape[-1]
    p = y_pred.detach().reshape(-1, n).float().contiguous()
    t = y_true.detach().reshape(-1, n).to(p.device, torch.float32).contiguous()
    ndcg, idcg = torch.empty(p.shape[0], device=p.device), torch.empty(p.shape[0], device=p.device)
    check(lib.unimm_neural_ndcg(ptr(p), ptr(t), p.shape[0], n, temperature, max_iter, tol, ptr(ndcg), ptr(idcg), _stream(p.device)))
    valid = idcg != 0
    loss = -(ndcg.sum() / valid.sum()) if bool(valid.any()) else torch.zeros((), device=p.device)
    return (loss, ndcg, idcg) if return_parts else loss


def neural_ndcg_loss_backward(y_pred: torch.Tensor, y_true: torch.Tensor, temperature: float = 1.0, max_iter: int = 50, tol: float = 1e-6,
                              grad_scale: float = 1.0):
    """-> (d loss / d y_pred * grad_scale  [same shape as y_pred], per-slate ndcg): what ``neuralNDCG_transposed(y_pred, y_true).backward()``
    leaves in ``y_pred.grad`` (dense_annotation_finetuning.py:288-296), computed by ``unimm_neural_ndcg_backward``."""
    if not y_pred.is_cuda:
        raise ValueError("neural_ndcg_loss_backward runs on the device the scores live on (CUDA)")
    n = y_pred.shape[-1]
    p = y_pred.detach().reshape(-1, n).float().contiguous()
    t = y_true.detach().reshape(-1, n).to(p.device, torch.float32).contiguous()
    d, ndcg = torch.empty_like(p), torch.empty(p.shape[0], device=p.device)
    cnt = torch.empty(1, dtype=torch.int32, device=p.device)
    check(lib.unimm_neural_ndcg_backward(ptr(p), ptr(t), p.shape[0], n, temperature, max_iter, tol, float(grad_scale), ptr(d), ptr(ndcg), ptr(cnt),
                                         _stream(p.device)))
    return d.view(*y_pred.shape), ndcg


def ensemble_normalise(probs: torch.Tensor) -> torch.Tensor:
    """probs [models, ..., options] -> [..., options]."""
    if not probs.is_cuda:
        raise ValueError("ensemble_normalise runs on the device the probabilities live on (CUDA)")
    m, n = probs.shape[0], probs.shape[-1]
    p = probs.detach().reshape(m, -1, n).float().contiguous()
    out = torch.empty(p.shape[1], n, device=p.device)
    check(lib.unimm_ensemble_normalise(ptr(p), m, p.shape[1], n, ptr(out), _stream(p.device)))
    return out.view(*probs.shape[1:])
