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
has ``prepare(step_items)`` / ``score(prepared, step_items)`` (``PackedScorer``) is pipelined: the next step is packed into
pinned staging (C++ packer, ``csrc/packer.cu``, straight from the per-image int64 tensors the reference's loader yields) on a
worker thread while the device scores the current one.
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
    arrays: object = None       # flat_packer.ImageArrays over the rounds' arrays (the loader's per-image tensors); built on demand

    def image_arrays(self):
        if self.arrays is None:
            from .flat_packer import ImageArrays
            self.arrays = ImageArrays.from_rounds(self.rounds, self.feat, self.loc, self.mask)
        return self.arrays


def synthetic_items(image_ids: Sequence[int], n_candidates: int = 100, only: Optional[Sequence[int]] = None, n_rounds: int = 10,
                    mode: str = "gen") -> List[DialogItem]:
    """The synthetic VisDial-shaped sweep of BASELINE.json configs[1] (seed = image id; gt option 0 as the reference's loader).

    ``only``: positions (into ``image_ids``) whose dialogs are actually generated — a rank passes its own shard; the other items
    carry just the ids / ground truth / relevance that the metrics need on every rank (``rounds`` is None there)."""
    from . import synthetic as syn
    out = []
    only = None if only is None else set(int(x) for x in only)
    for k, i in enumerate(image_ids):
        rng = np.random.RandomState(900001 + int(i))
        rel = rng.choice([0, 0, 0, 0.2, 0.4, 0.6, 0.8, 1.0], size=n_candidates).astype(np.float32)
        rel[0] = 1.0
        rel_round = int(rng.randint(n_rounds))
        if only is not None and k not in only:
            out.append(DialogItem(int(i), None, None, None, None, np.zeros(n_rounds, np.int64), rel_round, rel))
            continue
        (feat, loc, mask), rounds = syn.synth_dialog_rounds(int(i), rounds=tuple(range(1, n_rounds + 1)), n_candidates=n_candidates, mode=mode)
        it = DialogItem(int(i), feat, loc, mask, rounds, np.zeros(len(rounds), np.int64), rel_round, rel)
        if mode == "gen":
            it.image_arrays()
        out.append(it)
    return out


class PackedScorer:
    """Scores a step with ONE prefix-shared forward through the host-buffer C ABI (unimm_score_packed_host), in two phases so
    that ``run_sweep`` can overlap them: ``prepare`` (host only: the C++ packer fills a pinned staging buffer from the items'
    per-image int64 arrays, a few ms for 8 images) runs on a worker thread for step i + 1 while ``score`` (H2D + forward + D2H, one
    blocking C call that releases the GIL) runs step i.  ``depth`` + 1 packers rotate, so a prepared batch stays intact until it
    has been scored.

    Every step is checked for the precondition of the layout — the candidates of a round share their context token for token —
    unless ``verify_shared=False`` (the check is a threaded memcmp inside the packer, ~2 ms per step)."""

    def __init__(self, engine, verify_shared: bool = True, depth: int = 2, seq_len: int = 256, num_regions: int = 37,
                 feature_size: int = 2048, threads: int = 2):
        from .flat_packer import FlatPacker
        self.engine = engine
        self.verify_shared = verify_shared
        pinned = engine is not None
        if engine is not None:
            seq_len, num_regions, feature_size = engine.seq_len, engine.num_regions, engine.cfg.v_feature_size
            torch.cuda.set_device(engine.device)                 # pinned allocations belong to this rank's context
        self._packers = [FlatPacker(seq_len, num_regions, feature_size, pinned=pinned, threads=threads) for _ in range(depth + 1)]
        self._next = 0
        self._lock = threading.Lock()
        self._out = None

    def prepare(self, items: List[DialogItem]):
        if self.engine is not None and self.engine.device.type == "cuda":
            torch.cuda.set_device(self.engine.device)           # worker threads start on device 0: pin under this rank's context
        with self._lock:
            pk = self._packers[self._next % len(self._packers)]
            self._next += 1
        # the ranking reads the LM scores only (val_lm.py:124-139): scores-only layout
        return pk.pack([it.image_arrays() for it in items], scores_only=True, share_first_mask=True, verify_shared=self.verify_shared)

    def score(self, pb, items: List[DialogItem]) -> torch.Tensor:
        if self._out is None or self._out.numel() < pb.n_cands:
            self._out = torch.empty(pb.n_cands, dtype=torch.float32).pin_memory()
        out = self._out[:pb.n_cands]
        self.engine.score_packed_host(pb, out)
        return out.view(len(items), len(items[0].rounds), -1).clone()

    # ---- asynchronous halves (unimm_submit_packed_host / unimm_wait_packed): the device gets step i + 1 queued behind step i
    def submit(self, pb, items: List[DialogItem], slot: int) -> None:
        if not hasattr(self, "_outs"):
            self._outs = [None, None]
        if self._outs[slot] is None or self._outs[slot].numel() < pb.n_cands:
            self._outs[slot] = torch.empty(max(pb.n_cands, 1024), dtype=torch.float32).pin_memory()
        self.engine.submit_packed_host(pb, slot, self._outs[slot])
        self._shape = getattr(self, "_shape", {})
        self._shape[slot] = (len(items), len(items[0].rounds), pb.n_cands)

    def collect(self, slot: int) -> torch.Tensor:
        self.engine.wait_packed(slot)
        n_img, n_rounds, n_cands = self._shape[slot]
        return self._outs[slot][:n_cands].view(n_img, n_rounds, -1).clone()

    def __call__(self, items: List[DialogItem]) -> torch.Tensor:
        return self.score(self.prepare(items), items)


class NspScorer:
    """The discriminative ranking of the reference's val.py (:100-161): every option of every round is one dense sequence under the
    discriminative masks, its score is the NSP probability ``softmax(seq_relationship_score)[:, 0]``; with several models the
    per-round min-max-normalised probabilities are summed (``unimm_ensemble_normalise``, val.py:152-161).  ``engines``: one Engine per
    ensemble member on this rank's device.  Items must carry discriminative rounds (``synthetic_items(mode="dis")``)."""

    def __init__(self, engines, chunk: int = 250):
        self.engines = list(engines)
        self.chunk = min(chunk, min(e.max_sequences for e in self.engines))

    def __call__(self, items: List[DialogItem]) -> torch.Tensor:
        from .rank_loss import ensemble_normalise
        out = []
        for it in items:
            tok, seg, pos, desc = (torch.from_numpy(np.concatenate([getattr(r, f) for r in it.rounds])) for f in ("tokens", "segments", "positions", "desc"))
            n, n_rounds = tok.shape[0], len(it.rounds)
            feat, loc, mask = (torch.from_numpy(np.ascontiguousarray(a))[None] for a in (it.feat, it.loc, it.mask))
            index = torch.zeros(n, dtype=torch.int32)
            probs = []
            for eng in self.engines:
                p0 = []
                for s0 in range(0, n, self.chunk):
                    sl = slice(s0, s0 + self.chunk)
                    o = eng.forward(tok[sl], seg[sl], pos[sl], desc[sl], feat, loc, mask, feat_index=index[sl], want=("nsp_scores",))
                    p0.append(torch.softmax(o["nsp_scores"], 1)[:, 0])
                eng.check_ids()
                probs.append(torch.cat(p0).view(n_rounds, -1))
            out.append(ensemble_normalise(torch.stack(probs)).cpu())           # [rounds, options]
        return torch.stack(out)


def packed_scorer(engine, **kw) -> PackedScorer:
    return PackedScorer(engine, **kw)


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
              metrics_fn: Optional[Callable] = None, group=None, prefetch: int = 1, gather_device=None,
              timing: Optional[dict] = None, records_rank: Optional[int] = None) -> Dict[str, object]:
    """Score this rank's images, all-gather the scores, rank them and assemble metrics + EvalAI records (on every rank).

    ``items`` is the GLOBAL list (every rank passes the same one; only its own shard is scored).
    ``metrics_fn(scores [n,rounds,opts], gt_index [n,rounds], ndcg_scores [m,opts] | None, relevance [m,opts] | None)``
    returns a dict with ``ranks`` ([n, rounds, opts], 1-based) and the metric values.
    """
    n = len(items)
    mine = shard_units(n, rank, world)
    n_rounds = len(items[0].gt_index)
    local = []
    import time
    t_start = time.perf_counter()
    steps = [[items[i] for i in mine[s:s + images_per_step]] for s in range(0, len(mine), images_per_step)]

    def keep(out, step):
        if tuple(out.shape[:2]) != (len(step), n_rounds):
            raise ValueError("scorer must return [images, rounds, options]")
        local.append(out.float().cpu())

    if prefetch > 0 and hasattr(scorer, "submit") and hasattr(scorer, "collect") and steps:
        # asynchronous scorer: step i is packed (a few ms of host work) and queued on the device while step i - 1 still runs there;
        # its scores are collected only after step i + 1... has been queued, so the device never waits for the host
        pending = None
        for i, step in enumerate(steps):
            scorer.submit(scorer.prepare(step), step, i & 1)
            if pending is not None:
                keep(scorer.collect(pending[0]), pending[1])
            pending = (i & 1, step)
        keep(scorer.collect(pending[0]), pending[1])
    elif prefetch > 0 and hasattr(scorer, "prepare") and hasattr(scorer, "score") and steps:
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
    n_opt = local[0].shape[-1] if local else len(items[0].relevance)
    local_t = torch.cat(local).reshape(len(mine), n_rounds * n_opt) if local else torch.zeros(0, n_rounds * n_opt)
    t_scored = time.perf_counter()
    # the path's only exchange: with ``gather_device`` the local scores go up once and travel over NCCL / NVLink
    scores = gather_scores(local_t, n, rank, world, group, device=gather_device).view(n, n_rounds, n_opt)
    t_gathered = time.perf_counter()
    gt = torch.from_numpy(np.stack([it.gt_index for it in items])).long()
    ann = [i for i, it in enumerate(items) if it.relevance is not None and it.relevance_round >= 0]
    ndcg_scores = scores[torch.tensor(ann), torch.tensor([items[i].relevance_round for i in ann])] if ann else None
    rel = torch.from_numpy(np.stack([items[i].relevance for i in ann])).float() if ann else None
    m = dict(metrics_fn(scores, gt, ndcg_scores, rel))
    ranks = m.pop("ranks")
    t_metrics = time.perf_counter()
    # EvalAI records (val_lm.py:152-167) are written by ONE process: build them on ``records_rank`` only (None: every rank)
    records = None
    if records_rank is None or records_rank == rank:
        records = evalai_records(items, ranks)
    if timing is not None:
        timing.update(score_s=t_scored - t_start, gather_s=t_gathered - t_scored, metrics_s=t_metrics - t_gathered,
                      records_s=time.perf_counter() - t_metrics, steps=len(steps))
    return {"scores": scores.cpu(), "metrics": m, "predictions": records, "ranks": ranks}


def evalai_records(items: Sequence[DialogItem], ranks: torch.Tensor) -> List[dict]:
    """``{"image_id", "round_id" (1-based), "ranks"}`` per (image, round) — the reference's ranks_json (val_lm.py:152-167)."""
    ranks_l = ranks.tolist()
    return [{"image_id": int(it.image_id), "round_id": j + 1, "ranks": ranks_l[i][j]} for i, it in enumerate(items) for j in range(len(ranks_l[i]))]


def write_predictions(records: List[dict], path: str) -> None:
    """EvalAI submission file (val_lm.py:196-199)."""
    with open(path, "w") as f:
        json.dump(records, f)


def synthetic_sweep(n_images: int, images_per_step: int = 8, precision: str = "fp16", prefetch: int = 1, verify_shared: bool = True,
                    out_path: str = "", rank: int = 0, world: int = 1, local_rank: int = 0, on_timed_start=None) -> Dict[str, object]:
    """BASELINE.json configs[1]: ``n_images`` synthetic images x 10 rounds x 100 candidates, STRONG-scaled over ``world`` ranks
    (image i -> rank i mod world), scores all-gathered over NCCL, ranks / metrics on the GPU, EvalAI records on every rank.
    ``torch.distributed`` must already be initialised with the NCCL backend when ``world`` > 1.  Returns rank 0's report:
    wall clock of the whole sweep (first pack to last record; max over ranks) and its phases."""
    import time

    import torch.distributed as dist

    from .config import DEFAULT_CONFIG_PATH, ViLBertConfig
    from .engine import Engine
    from .sharding import shard_units
    from .weights import random_state_dict
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    cfg = ViLBertConfig.from_json_file(DEFAULT_CONFIG_PATH)
    mine = shard_units(n_images, rank, world)
    t0 = time.perf_counter()
    items = synthetic_items(range(n_images), only=mine)                     # generation is not part of the sweep's clock
    gen_s = time.perf_counter() - t0
    eng = Engine(cfg, random_state_dict(cfg, 0), precision=precision, max_sequences=images_per_step * 52, device=local_rank)
    scorer = packed_scorer(eng, verify_shared=verify_shared, depth=max(2, prefetch))
    metrics_fn = lambda s, g, ns, r: gpu_metrics(s, g, ns, r, dev)
    # warm-up: one step per rank through the whole driver (kernels, pinned staging, NCCL channels, metric kernels)
    warm = synthetic_items(range(world * images_per_step))
    run_sweep(warm, scorer, rank, world, images_per_step, metrics_fn=metrics_fn, gather_device=dev, prefetch=prefetch)
    del warm
    from .sharding import gather_scores
    gather_scores(torch.zeros(len(mine), len(items[0].gt_index) * len(items[0].relevance)), n_images, rank, world, device=dev)        # the exchange once at its real size
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    timing: Dict[str, float] = {}
    if on_timed_start is not None:
        on_timed_start()
    t0 = time.perf_counter()
    res = run_sweep(items, scorer, rank, world, images_per_step, metrics_fn=metrics_fn, gather_device=dev, prefetch=prefetch, timing=timing,
                    records_rank=0)
    torch.cuda.synchronize(dev)
    dt = torch.tensor([time.perf_counter() - t0, timing["score_s"]], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    n_cand = n_images * len(items[0].gt_index) * len(items[0].relevance)
    report = dict({k: (float(v) if not isinstance(v, (int, float)) else v) for k, v in res["metrics"].items()},
                  sweep_images=n_images, sweep_candidates=n_cand, sweep_seconds=float(dt[0]), sweep_candidates_per_sec=n_cand / float(dt[0]),
                  score_phase_seconds=float(dt[1]), score_phase_candidates_per_sec=n_cand / float(dt[1]),
                  phases_rank0={k: v for k, v in timing.items()}, prefetch=prefetch, images_per_step=images_per_step, world=world,
                  precision=precision, generation_seconds_rank0=gen_s, verify_shared=verify_shared,
                  note="wall clock of the whole driver (max over ranks): C++ packing from the per-image int64 arrays, H2D, forward, D2H, "
                       "NCCL all-gather of the scores, GPU ranks / metrics on every rank, EvalAI records on rank 0; synthetic generation excluded")
    if rank == 0 and out_path:
        write_predictions(res["predictions"], out_path)
    eng.close()
    return report


def main() -> None:
    import argparse
    import os

    import torch.distributed as dist
    ap = argparse.ArgumentParser(description="synthetic generative ranking sweep (one process per GPU under torchrun)")
    ap.add_argument("--images", type=int, default=16)
    ap.add_argument("--images-per-step", type=int, default=8)
    ap.add_argument("--precision", default="fp16")
    ap.add_argument("--out", default="")
    ap.add_argument("--no-verify", action="store_true", help="skip the per-step check that the candidates of a round share their context")
    ap.add_argument("--prefetch", type=int, default=1, help="0: pack and score serially; n: pack steps i+1..i+n on n worker threads while step i is scored")
    a = ap.parse_args()
    world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", "1"), ("RANK", "0"), ("LOCAL_RANK", "0")))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rep = synthetic_sweep(a.images, a.images_per_step, a.precision, a.prefetch, not a.no_verify, a.out, rank, world, local)
    if rank == 0:
        print(json.dumps(rep))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
