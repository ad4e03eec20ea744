"""CPU, world_size 2 over gloo: unit partition + the single score all-gather of the multi-GPU path."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from unimm_b200.sharding import gather_scores, shard_units


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _score_of(unit, n_opt):          # a deterministic stand-in for "the scores rank r computed for unit u"
    return torch.arange(n_opt, dtype=torch.float32) * 0.01 + unit


def _worker(rank, world, port, n_units, n_opt, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        mine = shard_units(n_units, rank, world)
        local = torch.stack([_score_of(u, n_opt) for u in mine]) if mine else torch.zeros(0, n_opt)
        full = gather_scores(local, n_units)
        want = torch.stack([_score_of(u, n_opt) for u in range(n_units)])
        q.put((rank, bool(torch.equal(full, want)), len(mine)))
    finally:
        dist.destroy_process_group()


def _run(n_units, world=2, n_opt=100):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_units, n_opt, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    return sorted(res)


def test_partition_is_a_balanced_cover():
    for n, w in ((20640, 8), (7, 2), (1, 4), (10, 3)):
        parts = [shard_units(n, r, w) for r in range(w)]
        assert sorted(sum(parts, [])) == list(range(n))
        assert max(map(len, parts)) - min(map(len, parts)) <= 1


def test_all_gather_of_scores_world2_even():
    res = _run(10)
    assert all(ok for _, ok, _ in res) and [n for _, _, n in res] == [5, 5]


def test_all_gather_of_scores_world2_ragged():
    res = _run(7)
    assert all(ok for _, ok, _ in res) and [n for _, _, n in res] == [4, 3]


def test_single_rank_is_identity():
    x = torch.randn(3, 100)
    assert torch.equal(gather_scores(x, 3, rank=0, world=1), x)
