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


def lm_head_block_backward(x: torch.Tensor, params: dict, labels: torch.Tensor, weight: torch.Tensor, grad_scale: float = 1.0,
                           precision: str = "fp16"):
    """Backward of the WHOLE masked-LM head of the reference (``cls.predictions``: transform dense -> erf-GELU -> LayerNorm -> tied
    decoder + bias, models/vilbert_dialog.py:982-986, :1023-1026) under the likelihood / unlikelihood loss (:1577-1595), on the rows
    ``x [n, 768]`` that carry a label: chains ``unimm_k_lm_head_backward`` (decoder + loss), ``unimm_k_layernorm_backward``,
    ``unimm_k_gelu_backward`` and ``unimm_k_linear_backward`` (transform).  ``params``: ``transform.dense.weight/bias``,
    ``transform.LayerNorm.weight/bias``, ``decoder.weight`` (= word embeddings), ``bias``.  Returns the gradient with respect to ``x``
    (what the encoder's backward starts from) and to every parameter of the head, fp32."""
    if not x.is_cuda:
        raise ValueError("lm_head_block_backward runs on a CUDA device (B200); there is no CPU path")
    dev = x.device
    dt = torch.float16 if precision == "fp16" else torch.bfloat16
    kind = LP_FP16 if precision == "fp16" else LP_BF16
    P = {k: v.to(dev, torch.float32).contiguous() for k, v in params.items()}
    n, K = x.shape
    stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    x16 = x.to(dt).contiguous()
    Wt16 = P["transform.dense.weight"].to(dt).contiguous()
    # forward recompute of the head's activations (the engine's forward keeps none of them): t = x Wt^T + bt, g = gelu(t), h = LN(g)
    t = torch.empty(n, K, device=dev)
    check(lib.unimm_k_gemm_lp(ptr(x16), K, ptr(Wt16), K, n, K, K, ptr(P["transform.dense.bias"]), None, 0, 0, ptr(t), K, None, 0, 0, 0, kind, stream))
    g = torch.empty(n, K, device=dev)
    check(lib.unimm_k_gemm_lp(ptr(x16), K, ptr(Wt16), K, n, K, K, ptr(P["transform.dense.bias"]), None, 0, 1, ptr(g), K, None, 0, 0, 0, kind, stream))
    h16 = torch.empty(n, K, device=dev, dtype=dt)
    check(lib.unimm_k_layernorm(ptr(g), K, n, K, ptr(P["transform.LayerNorm.weight"]), ptr(P["transform.LayerNorm.bias"]), None, ptr(h16), kind, stream))
    # decoder + loss
    head = lm_head_backward(h16, P["decoder.weight"], P["bias"], labels, weight, grad_scale=grad_scale, precision=precision)
    # LayerNorm, GELU, transform
    dg, dgamma, dbeta = torch.empty(n, K, device=dev), torch.empty(K, device=dev), torch.empty(K, device=dev)
    check(lib.unimm_k_layernorm_backward(ptr(head["dH"]), ptr(g), n, K, ptr(P["transform.LayerNorm.weight"]), ptr(dg), ptr(dgamma), ptr(dbeta), stream))
    check(lib.unimm_k_gelu_backward(ptr(dg), ptr(t), n * K, ptr(dg), stream))
    lin = linear_backward(dg, x16, Wt16, precision=precision)
    return {"dx": lin["dX"], "transform.dense.weight": lin["dW"], "transform.dense.bias": lin["db"], "transform.LayerNorm.weight": dgamma,
            "transform.LayerNorm.bias": dbeta, "decoder.weight": head["dE"], "bias": head["dbias"], "logp": head["logp"]}
