"""VisDial ranking metrics on the GPU (reference utils/visdial_metrics.py:21-193) over the score tensor.

``rank_metrics(scores, gt_index, relevance)`` returns the reference's metric names (r@1, r@5, r@10, mean, mrr, ndcg)
plus ``ties`` (number of exactly tied score pairs: the reference's unstable sort orders those arbitrarily) and the
1-based ``ranks``.  NDCG follows NDCG.observe: one relevance row per score row, k = number of non-zero relevances.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional

import torch

from ._lib import check, lib, ptr


def rank_metrics(scores: torch.Tensor, gt_index: Optional[torch.Tensor] = None, relevance: Optional[torch.Tensor] = None,
                 return_ranks: bool = True) -> Dict[str, object]:
    if not scores.is_cuda:
        raise ValueError("rank_metrics runs on the device the scores live on (CUDA)")
    n_opt = scores.shape[-1]
    s = scores.detach().reshape(-1, n_opt).to(torch.float32).contiguous()
    rows = s.shape[0]
    dev = s.device
    gt = None if gt_index is None else gt_index.reshape(-1).to(device=dev, dtype=torch.int32).contiguous()
    rel = None if relevance is None else relevance.reshape(-1, n_opt).to(device=dev, dtype=torch.float32).contiguous()
    if gt is not None and gt.numel() != rows:
        raise ValueError("gt_index must have one entry per score row")
    if rel is not None and rel.shape[0] != rows:
        raise ValueError("relevance must have one row per score row")
    ranks = torch.empty(rows, n_opt, dtype=torch.int32, device=dev) if return_ranks else None
    sums = torch.zeros(9, dtype=torch.float64, device=dev)
    stream = torch.cuda.current_stream(dev).cuda_stream
    check(lib.unimm_rank_metrics(ptr(s), rows, n_opt, ptr(gt), ptr(rel), ptr(ranks), ptr(sums), C.c_void_p(stream)))
    v = sums.cpu().tolist()
    out: Dict[str, object] = {"ties": int(v[8])}
    if gt is not None:
        n = v[0]
        out.update({"r@1": v[1] / n, "r@5": v[2] / n, "r@10": v[3] / n, "mean": v[4] / n, "mrr": v[5] / n})
    if rel is not None and v[7] > 0:
        out["ndcg"] = v[6] / v[7]
    if return_ranks:
        out["ranks"] = ranks.view(*scores.shape)
    return out
