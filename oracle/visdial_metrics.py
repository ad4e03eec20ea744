"""ORACLE (test infrastructure): restatement of the reference's ranking metrics.

Follows ``/root/reference/utils/visdial_metrics.py``: ``scores_to_ranks`` (:21-39), ``SparseGTMetrics``
(:41-115) and ``NDCG`` (:117-193).  Vectorised instead of the reference's Python loops; pinned by
``tests/golden/make_golden.py`` (equal ranks / metrics to the reference classes on the same scores).
"""
from __future__ import annotations

import torch


def scores_to_ranks(scores: torch.Tensor) -> torch.Tensor:
    """1-based rank of each option, highest score → rank 1 (visdial_metrics.py:21-39)."""
    shape = scores.shape
    flat = scores.reshape(-1, shape[-1])
    order = flat.sort(1, descending=True)[1]
    ranks = torch.empty_like(order)
    ranks.scatter_(1, order, torch.arange(1, shape[-1] + 1).expand_as(order))
    return ranks.view(shape)


def sparse_metrics(scores: torch.Tensor, gt_index: torch.Tensor) -> dict:
    """R@1/5/10, mean rank, MRR over [batch, rounds, options] scores (visdial_metrics.py:52-90)."""
    ranks = scores_to_ranks(scores)
    gt = ranks.reshape(-1, ranks.shape[-1]).gather(1, gt_index.reshape(-1, 1).long())[:, 0].float()
    return {"r@1": (gt <= 1).float().mean().item(), "r@5": (gt <= 5).float().mean().item(),
            "r@10": (gt <= 10).float().mean().item(), "mean": gt.mean().item(), "mrr": gt.reciprocal().mean().item()}


def ndcg(scores: torch.Tensor, relevance: torch.Tensor) -> float:
    """NDCG over [batch, options] (visdial_metrics.py:122-176): k = #non-zero relevance per row."""
    ranks = scores_to_ranks(scores.unsqueeze(1)).squeeze(1)
    rankings = ranks.sort(-1)[1]
    best = relevance.sort(-1, descending=True)[1]
    vals = []
    for b in range(scores.shape[0]):
        k = int((relevance[b] != 0).sum())
        disc = torch.log2(torch.arange(k).float() + 2)
        dcg = (relevance[b][rankings[b][:k]].float() / disc).sum()
        ideal = (relevance[b][best[b][:k]].float() / disc).sum()
        vals.append(dcg / ideal)
    return float(sum(vals) / len(vals))
