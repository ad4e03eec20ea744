"""Generative ranking sweep: the caller of the hot path (reference val_lm.py:48-200).

For every image the 10 dialog rounds x 100 candidate answers are scored by sequence log-likelihood
(val_lm.py:104-139), the scores are ranked (utils/visdial_metrics.py:21-39), sparse metrics (R@k, mean rank, MRR)
use the ground-truth option and NDCG uses the dense relevance of one annotated round per image
(val_lm.py:169-178), and the ranks are written in the EvalAI format (val_lm.py:152-167:
``{"image_id", "round_id" (1-based), "ranks"}`` per round).

What differs from the reference loop — by design, not in results:
  * images are sharded over ranks (one process per GPU, ``image i -> rank i mod N``; all 10 rounds of an image stay
    on one rank so they share its feature block) instead of nn.DataParallel scattering every chunk;
  * a step packs ``images_per_step`` images into ONE prefix-shared forward (unimm_b200.packing) instead of 40
    chunks of 25 dense sequences per image;
  * only the ``[images, 10, 100]`` score tensor is exchanged (one all-gather), ranks / metrics run on the GPU
    (csrc/metrics.cu) or, without one, through any ``metrics_fn`` the caller passes.

``scorer(step_items) -> float32 [len(step_items), rounds, options]`` is the only device-touching piece, so the
sharding / gathering / bookkeeping is tested on CPU with a stand-in scorer (tests/test_val_sweep_cpu.py).  A scorer that also
has ``prepare(step_items)`` / ``score(prepared, step_items)`` (``PackedScorer``) is pipelined: the next step is packed and
pinned on a worker thread while the device scores the current one.
"""
from __future__ import annotations

import json
import threading
from dataclasses import dataclass
from typing import Callable, Dict, List, Optional, Sequence

import numpy as np
import torch

from .sharding import gather_scores, shard_units


@dataclass
class DialogItem:
    """One image of the sweep: its feature block and its dialog rounds (``unimm_b200.synthetic.Round``-like units)."""
    image_id: int
    feat: np.ndarray            # [R, 2048]
    loc: np.ndarray             # [R, 5]
    mask: np.ndarray            # [R]
    rounds: list                # n_rounds units of n_options candidates
    gt_index: np.ndarray        # [n_rounds] ground-truth option of every round (dataloader_visdial.py:337-342)
    relevance_round: int = -1   # 0-based round that carries dense annotations (-1: none)
    relevance: Optional[np.ndarray] = None   # [n_options]


def synthetic_items(image_ids: Sequence[int], n_candidates: int = 100) -> List[DialogItem]:
    """The synthetic VisDial-shaped sweep of BASELINE.json configs[1] (seed = image id; gt option 0 as the reference's loader)."""
    from . import synthetic as syn
    out = []
    for i in image_ids:
        (feat, loc, mask), rounds = syn.synth_dialog_rounds(int(i), n_candidates=n_candidates)
        rng = np.random.RandomState(900001 + int(i))
        rel = rng.choice([0, 0, 0, 0.2, 0.4, 0.6, 0.8, 1.0], size=n_candidates).astype(np.float32)
        rel[0] = 1.0
        out.append(DialogItem(int(i), feat, loc, mask, rounds, np.zeros(len(rounds), np.int64), int(rng.randint(len(rounds))), rel))
    return out


class PackedScorer:
    """Scores a step with ONE prefix-shared forward through the host-buffer C ABI (unimm_score_packed_host), in two phases so
    that ``run_sweep`` can overlap them: ``prepare`` (host only: pack + pin, ~20-30 ms for 8 images) runs on a worker thread for
    step i + 1 while ``score`` (H2D + forward + D2H, one blocking C call that releases the GIL) runs step i.

    The packer's context-equality check is made on the first ``verify_steps`` steps of a scorer's life only: it catches a loader
    whose candidates do not share their context, and costs a third of the packing time."""

    def __init__(self, engine, verify_steps: int = 1):
        self.engine = engine
        self.verify_steps = verify_steps
        self._prepared = 0
        self._lock = threading.Lock()

    def prepare(self, items: List[DialogItem]):
        from .packing import pack_units, units_from_rounds
        if self.engine is not None and self.engine.device.type == "cuda":
            torch.cuda.set_device(self.engine.device)           # worker threads start on device 0: pin under this rank's context
        rounds, slots = [], []
        for s, it in enumerate(items):
            rounds += list(it.rounds)
            slots += [s] * len(it.rounds)
        with self._lock:
            verify = self._prepared < self.verify_steps
            self._prepared += 1
        # the ranking reads the LM scores only (val_lm.py:124-139): scores-only layout
        pb = pack_units(units_from_rounds(rounds, slots), np.stack([it.feat for it in items]), np.stack([it.loc for it in items]),
                        np.stack([it.mask for it in items]), scores_only=True, verify_shared=verify)
        return pb.pin() if torch.cuda.is_available() else pb

    def score(self, pb, items: List[DialogItem]) -> torch.Tensor:
        out = torch.empty(pb.n_cands, dtype=torch.float32).pin_memory()
        self.engine.score_packed_host(pb, out)
        return out.view(len(items), len(items[0].rounds), -1).clone()

    def __call__(self, items: List[DialogItem]) -> torch.Tensor:
        return self.score(self.prepare(items), items)


def packed_scorer(engine) -> PackedScorer:
    return PackedScorer(engine)


def gpu_metrics(scores: torch.Tensor, gt_index: torch.Tensor, ndcg_scores: Optional[torch.Tensor], relevance: Optional[torch.Tensor],
                device) -> Dict[str, object]:
    from .metrics import rank_metrics
    m = rank_metrics(scores.to(device), gt_index)
    ranks = m.pop("ranks").cpu()
    if ndcg_scores is not None and ndcg_scores.shape[0] > 0:
        m["ndcg"] = rank_metrics(ndcg_scores.to(device), relevance=relevance, return_ranks=False)["ndcg"]
    m["ranks"] = ranks
    return m


def run_sweep(items: Sequence[DialogItem], scorer, rank: int = 0, world: int = 1, images_per_step: int = 8,
              metrics_fn: Optional[Callable] = None, group=None, prefetch: int = 1) -> Dict[str, object]:
    """Score this rank's images, all-gather the scores, rank them and assemble metrics + EvalAI records (on every rank).

    ``items`` is the GLOBAL list (every rank passes the same one; only its own shard is scored).
    ``metrics_fn(scores [n,rounds,opts], gt_index [n,rounds], ndcg_scores [m,opts] | None, relevance [m,opts] | None)``
    returns a dict with ``ranks`` ([n, rounds, opts], 1-based) and the metric values.
    """
    n = len(items)
    mine = shard_units(n, rank, world)
    n_rounds = len(items[0].rounds)
    local = []
    steps = [[items[i] for i in mine[s:s + images_per_step]] for s in range(0, len(mine), images_per_step)]

    def keep(out, step):
        if tuple(out.shape[:2]) != (len(step), n_rounds):
            raise ValueError("scorer must return [images, rounds, options]")
        local.append(out.float().cpu())

    if prefetch > 0 and hasattr(scorer, "prepare") and hasattr(scorer, "score") and steps:
        # two-phase scorer: the host-side preparation of step i + 1 overlaps the device work of step i
        # (``prefetch`` steps in flight on as many worker threads: the packer spends most of its time inside numpy, GIL released)
        from collections import deque
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(max_workers=prefetch) as pool:
            futs = deque(pool.submit(scorer.prepare, st) for st in steps[:prefetch])
            for i, step in enumerate(steps):
                prepared = futs.popleft().result()
                if i + prefetch < len(steps):
                    futs.append(pool.submit(scorer.prepare, steps[i + prefetch]))
                keep(scorer.score(prepared, step), step)
    else:
        for step in steps:
            keep(scorer(step), step)
    n_opt = local[0].shape[-1] if local else len(items[0].rounds[0].tokens)
    local_t = torch.cat(local).reshape(len(mine), n_rounds * n_opt) if local else torch.zeros(0, n_rounds * n_opt)
    scores = gather_scores(local_t, n, rank, world, group).view(n, n_rounds, n_opt)       # the path's only exchange
    gt = torch.from_numpy(np.stack([it.gt_index for it in items])).long()
    ann = [i for i, it in enumerate(items) if it.relevance is not None and it.relevance_round >= 0]
    ndcg_scores = torch.stack([scores[i, items[i].relevance_round] for i in ann]) if ann else None
    rel = torch.from_numpy(np.stack([items[i].relevance for i in ann])).float() if ann else None
    m = dict(metrics_fn(scores, gt, ndcg_scores, rel))
    ranks = m.pop("ranks")
    records = [{"image_id": int(items[i].image_id), "round_id": j + 1, "ranks": [int(r) for r in ranks[i, j].tolist()]}
               for i in range(n) for j in range(n_rounds)]
    return {"scores": scores, "metrics": m, "predictions": records}


def write_predictions(records: List[dict], path: str) -> None:
    """EvalAI submission file (val_lm.py:196-199)."""
    with open(path, "w") as f:
        json.dump(records, f)


def main() -> None:
    import argparse
    import os

    import torch.distributed as dist

    from .config import DEFAULT_CONFIG_PATH, ViLBertConfig
    from .engine import Engine
    from .weights import random_state_dict
    ap = argparse.ArgumentParser(description="synthetic generative ranking sweep (one process per GPU under torchrun)")
    ap.add_argument("--images", type=int, default=16)
    ap.add_argument("--images-per-step", type=int, default=8)
    ap.add_argument("--precision", default="fp16")
    ap.add_argument("--out", default="")
    ap.add_argument("--prefetch", type=int, default=1, help="0: pack and score serially; n: pack steps i+1..i+n on n worker threads while step i is scored")
    a = ap.parse_args()
    world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", "1"), ("RANK", "0"), ("LOCAL_RANK", "0")))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    cfg = ViLBertConfig.from_json_file(DEFAULT_CONFIG_PATH)
    items = synthetic_items(range(a.images))
    eng = Engine(cfg, random_state_dict(cfg, 0), precision=a.precision, max_sequences=a.images_per_step * 52, device=local)
    scorer = packed_scorer(eng)
    # the score tensor is exchanged as a host tensor: 8 MB for the full val sweep, a gloo group next to NCCL is plenty
    group = dist.new_group(backend="gloo") if world > 1 else None
    import time
    metrics_fn = lambda s, g, ns, r: gpu_metrics(s, g, ns, r, dev)
    run_sweep(items[:world * a.images_per_step], scorer, rank, world, a.images_per_step, metrics_fn=metrics_fn, group=group)     # warm-up step
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    res = run_sweep(items, scorer, rank, world, a.images_per_step, metrics_fn=metrics_fn, group=group, prefetch=a.prefetch)
    torch.cuda.synchronize(dev)
    dt = time.perf_counter() - t0
    if rank == 0:
        n_cand = sum(len(r.tokens) for it in items for r in it.rounds)
        print(json.dumps(dict({k: v for k, v in res["metrics"].items()}, sweep_candidates=n_cand, sweep_seconds=dt,
                              sweep_candidates_per_sec=n_cand / dt, prefetch=a.prefetch,
                              note="whole driver incl. host-side packing, pinning, score gather, ranks and metrics; wall clock")))
        if a.out:
            write_predictions(res["predictions"], a.out)
    eng.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
