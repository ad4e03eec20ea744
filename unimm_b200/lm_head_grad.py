"""Backward of the fused LM head + likelihood / unlikelihood loss on the device (``unimm_k_lm_head_backward``): the first piece of
SURVEY.md §8f item 1 (reference models/vilbert_dialog.py:1577-1595 under the backward of train.py:453-463).

Given the hidden states of the labelled rows ``h [n, 768]`` (output of the LM transform + LayerNorm), the tied decoder
``E [30522, 768]``, its bias, the labels and the token weights, returns the gradients of ``grad_scale * sum_i loss_i`` with respect
to ``h``, ``E`` and the bias, where ``loss_i = -w_i log p_i(y_i)`` for likelihood rows (``w_i > 0``) and
``-log(max(1 - p_i(y_i), 1e-6))`` for unlikelihood rows (``w_i == -1``).  The reference's ``masked_lm_loss`` is this sum with
``grad_scale = 1 / #(w != 0)``.  No ``[n, 30522]`` fp32 logits exist at any point.
"""
from __future__ import annotations

import ctypes as C

import torch

from ._lib import LP_BF16, LP_FP16, check, lib, ptr


def lm_head_backward(h: torch.Tensor, E: torch.Tensor, bias: torch.Tensor, labels: torch.Tensor, weight: torch.Tensor,
                     grad_scale: float = 1.0, precision: str = "fp16"):
    """h [n,K] and E [V,K] fp32 (cast here to the 16-bit operand format) or already 16-bit; -> dict(dH, dE, dbias, logp)."""
    if not h.is_cuda:
        raise ValueError("lm_head_backward runs on a CUDA device (B200); there is no CPU path")
    dt = torch.float16 if precision == "fp16" else torch.bfloat16
    kind = LP_FP16 if precision == "fp16" else LP_BF16
    dev = h.device
    h16, E16 = h.to(dt).contiguous(), E.to(dt).contiguous()
    n, K = h16.shape
    V = E16.shape[0]
    nbytes = lib.unimm_k_lm_head_backward_scratch(n, V, K)
    scratch = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    dH, dE = torch.empty(n, K, device=dev), torch.empty(V, K, device=dev)
    dbias, logp = torch.empty(V, device=dev), torch.empty(n, device=dev)
    lab = labels.to(dev, torch.int32).contiguous()
    w = weight.to(dev, torch.float32).contiguous()
    b = bias.to(dev, torch.float32).contiguous()
    stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    check(lib.unimm_k_lm_head_backward(ptr(h16), K, ptr(E16), K, n, V, K, ptr(b), ptr(lab), ptr(w), float(grad_scale), ptr(dH), ptr(dE),
                                       ptr(dbias), ptr(logp), ptr(scratch), nbytes, kind, stream))
    return {"dH": dH, "dE": dE, "dbias": dbias, "logp": logp}


def linear_backward(dY: torch.Tensor, X: torch.Tensor, W: torch.Tensor, precision: str = "fp16", want=("dX", "dW", "db")):
    """dgrad / wgrad / bias gradient of ``y = x W^T + b`` on tcgen05 (``unimm_k_linear_backward``): dY fp32 [M,N], X [M,K], W [N,K]
    (cast here to the 16-bit operand format) -> dict of fp32 tensors."""
    if not dY.is_cuda:
        raise ValueError("linear_backward runs on a CUDA device (B200); there is no CPU path")
    dt = torch.float16 if precision == "fp16" else torch.bfloat16
    kind = LP_FP16 if precision == "fp16" else LP_BF16
    dev = dY.device
    dY = dY.float().contiguous()
    X16, W16 = X.to(dev, dt).contiguous(), W.to(dev, dt).contiguous()
    M, N = dY.shape
    K = X16.shape[1]
    nbytes = lib.unimm_k_linear_backward_scratch(M, N, K)
    scratch = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    out = {}
    if "dX" in want:
        out["dX"] = torch.empty(M, K, device=dev)
    if "dW" in want:
        out["dW"] = torch.empty(N, K, device=dev)
    if "db" in want:
        out["db"] = torch.empty(N, device=dev)
    stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    check(lib.unimm_k_linear_backward(ptr(dY), N, ptr(X16), K, ptr(W16), K, M, N, K, ptr(out.get("dX")), ptr(out.get("dW")), ptr(out.get("db")),
                                      ptr(scratch), nbytes, kind, stream))
    return out
