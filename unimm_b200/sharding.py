"""Multi-GPU partitioning of the generative-scoring sweep.

Units (image, round) are independent, so the path shards with NO collective on the data path: one process per GPU
(torchrun), weights replicated, unit ``u`` -> rank ``u mod N`` (rounds of an image differ in context length, so
interleaving balances the ranks), and ONE exchange at the end — an all-gather of the ``[units,100]`` score tensor
so that rank 0 (or every rank) can compute MRR / R@k / NDCG.  This replaces the reference's per-forward
``nn.DataParallel`` scatter / replicate / gather (utils/data_parallel.py:91-132, val_lm.py:253-257).
"""
from __future__ import annotations

from typing import List, Optional

import torch
import torch.distributed as dist


def shard_units(n_units: int, rank: int, world: int) -> List[int]:
    """Unit ids owned by ``rank``."""
    return list(range(rank, n_units, world))


def gather_scores(local_scores: torch.Tensor, n_units: int, rank: Optional[int] = None, world: Optional[int] = None,
                  group=None, device=None) -> torch.Tensor:
    """All-gather per-unit score rows into the global ``[n_units, n_options]`` tensor (every rank gets it).

    ``local_scores[i]`` belongs to unit ``shard_units(n_units, rank, world)[i]``.  Ranks may own different numbers of
    units; rows are padded to the maximum count for the collective and dropped afterwards.  ``device``: run the collective on
    that device (NCCL over NVLink: one ncclAllGather of ``[ceil(n/world), n_options]`` per rank) and return a tensor there;
    None: on the tensor's own device (gloo for the CPU tests).
    """
    if world is None:
        world = dist.get_world_size(group) if dist.is_initialized() else 1
    if rank is None:
        rank = dist.get_rank(group) if dist.is_initialized() else 0
    mine = shard_units(n_units, rank, world)
    if local_scores.shape[0] != len(mine):
        raise ValueError(f"rank {rank} owns {len(mine)} units but passed {local_scores.shape[0]} score rows")
    n_opt = local_scores.shape[1]
    if device is not None:
        local_scores = local_scores.to(device, non_blocking=True)
    if world == 1:
        return local_scores.clone()
    per_rank = (n_units + world - 1) // world
    padded = local_scores.new_zeros(per_rank, n_opt)
    padded[: len(mine)] = local_scores
    out = local_scores.new_empty(world * per_rank, n_opt)
    dist.all_gather_into_tensor(out, padded, group=group)
    # unit u lives at row (u mod world) * per_rank + u // world of the gathered tensor
    u = torch.arange(n_units, device=out.device)
    return out[(u % world) * per_rank + u // world]
