"""CPU, world_size 2 over gloo: the generative ranking sweep's sharding, score gather, metrics and EvalAI records
(unimm_b200/val_sweep.py), with a stand-in scorer and the oracle's metric functions."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import visdial_metrics as om
from unimm_b200.val_sweep import DialogItem, run_sweep


class _Unit:            # the only thing run_sweep needs from a round without a scorer that reads it
    def __init__(self, n):
        self.tokens = np.zeros((n, 4), np.int64)


def _items(n_images, n_rounds=3, n_opt=12):
    out = []
    for i in range(n_images):
        rng = np.random.RandomState(50 + i)
        rel = rng.choice([0, 0, 0.5, 1.0], size=n_opt).astype(np.float32)
        rel[0] = 1.0
        out.append(DialogItem(1000 + i, np.zeros((2, 4), np.float32), np.zeros((2, 5), np.float32), np.ones(2, np.float32),
                              [_Unit(n_opt) for _ in range(n_rounds)], rng.randint(0, n_opt, size=n_rounds), int(rng.randint(n_rounds)),
                              rel))
    return out


def _scorer(step):      # deterministic in the image id only
    return torch.stack([torch.from_numpy(np.random.RandomState(it.image_id).randn(len(it.rounds), len(it.rounds[0].tokens)).astype(np.float32))
                        for it in step])


def _metrics(scores, gt, ndcg_scores, rel):
    m = dict(om.sparse_metrics(scores, gt))
    m["ranks"] = om.scores_to_ranks(scores)
    if ndcg_scores is not None:
        m["ndcg"] = om.ndcg(ndcg_scores, rel)
    return m


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_images, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        res = run_sweep(_items(n_images), _scorer, rank, world, images_per_step=2, metrics_fn=_metrics)
        q.put((rank, res["scores"].numpy(), {k: float(v) for k, v in res["metrics"].items()}, res["predictions"]))
    finally:
        dist.destroy_process_group()


def test_sweep_world2_equals_single_process():
    n_images = 5                                                  # ragged: rank 0 owns 3 images, rank 1 owns 2
    single = run_sweep(_items(n_images), _scorer, 0, 1, images_per_step=2, metrics_fn=_metrics)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_images, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=180) for _ in procs], key=lambda r: r[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, scores, metrics, preds in res:
        np.testing.assert_array_equal(scores, single["scores"].numpy())
        assert preds == single["predictions"]
        for k, v in single["metrics"].items():
            assert abs(metrics[k] - float(v)) < 1e-12, k


def test_prediction_records_are_evalai_shaped():
    res = run_sweep(_items(3), _scorer, 0, 1, images_per_step=8, metrics_fn=_metrics)
    recs = res["predictions"]
    assert len(recs) == 3 * 3
    assert [r["round_id"] for r in recs[:3]] == [1, 2, 3] and recs[0]["image_id"] == 1000
    for r in recs:
        assert sorted(r["ranks"]) == list(range(1, 13))           # 1-based ranks, a permutation (val_lm.py:152-167)
    # rank 1 = the best-scoring option
    s = res["scores"]
    assert recs[4]["ranks"][int(s[1, 1].argmax())] == 1
    assert set(res["metrics"]) >= {"r@1", "r@5", "r@10", "mean", "mrr", "ndcg"}


class _TwoPhase:
    """Stand-in for PackedScorer: prepare() is slow host work, score() the 'device' call; both log when they start and end."""
    def __init__(self):
        import threading
        self.log, self.lock, self.threads = [], threading.Lock(), set()

    def _mark(self, what, step):
        import threading
        import time
        with self.lock:
            self.log.append((what, step[0].image_id, time.perf_counter()))
            if what.startswith("prepare"):
                self.threads.add(threading.get_ident())

    def prepare(self, step):
        import time
        self._mark("prepare_begin", step)
        time.sleep(0.05)
        self._mark("prepare_end", step)
        return [it.image_id for it in step]

    def score(self, prepared, step):
        import time
        assert prepared == [it.image_id for it in step]            # every step is scored with ITS preparation
        self._mark("score_begin", step)
        time.sleep(0.05)
        out = _scorer(step)
        self._mark("score_end", step)
        return out

    def __call__(self, step):
        return self.score(self.prepare(step), step)


def test_two_phase_scorer_is_pipelined_and_gives_the_same_result():
    import threading
    items = _items(7)
    serial = run_sweep(items, _scorer, 0, 1, images_per_step=2, metrics_fn=_metrics)
    tp = _TwoPhase()
    piped = run_sweep(items, tp, 0, 1, images_per_step=2, metrics_fn=_metrics)
    np.testing.assert_array_equal(piped["scores"].numpy(), serial["scores"].numpy())
    assert piped["predictions"] == serial["predictions"]
    assert tp.threads and threading.get_ident() not in tp.threads                      # preparation ran on the worker thread
    t = {(w, i): ts for w, i, ts in tp.log}
    firsts = sorted({i for _, i, _ in tp.log})
    for a, b in zip(firsts, firsts[1:]):                                               # step b is prepared while step a is scored
        assert t[("prepare_begin", b)] < t[("score_end", a)]
    off = _TwoPhase()
    plain = run_sweep(items, off, 0, 1, images_per_step=2, metrics_fn=_metrics, prefetch=0)
    np.testing.assert_array_equal(plain["scores"].numpy(), serial["scores"].numpy())
    assert off.threads == {threading.get_ident()}
