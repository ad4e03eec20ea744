"""CPU, world_size 2 over gloo: the data-parallel training step (one process per GPU, ONE all-reduce of the flat gradient buffer,
1 / world folded into AdamW) equals the mean of the replicas' gradients and leaves identical parameters on every rank.  The torch-fp64
operations of tests/torch_train_ops.py stand in for the CUDA kernels (there is no GPU here); the collective plumbing is the product's."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _slice(batch, lo, hi):
    per_seq = ("tokens", "segments", "positions", "labels", "weights", "desc", "next_sentence_label", "seq_image")
    return {k: (v[lo:hi] if k in per_seq else v) for k, v in batch.items()}


def _worker(rank, world, port, q):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.set_num_threads(2)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from test_train_step_cpu import _train_inputs
        from torch_train_ops import TorchOps
        from unimm_b200.config import tiny_config
        from unimm_b200.train_step import TrainStep
        from unimm_b200.weights import random_state_dict
        cfg = tiny_config()
        sd = random_state_dict(cfg, seed=5, perturbed=True)
        _, _, batch = _train_inputs(4)
        ts = TrainStep(cfg, sd, TorchOps(), lr=1e-3, image_lr=1e-3, warmup_steps=0)
        assert ts.world == world
        vals = ts.step(_slice(batch, 2 * rank, 2 * rank + 2))
        g = ts.params.g.clone()
        p = ts.params.p.clone()
        q.put((rank, vals, g.numpy(), p.numpy()))
    finally:
        dist.destroy_process_group()


@pytest.mark.slow
def test_data_parallel_step_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=600) for _ in procs], key=lambda r: r[0])
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    (_, v0, g0, p0), (_, v1, g1, p1) = res
    assert (g0 == g1).all() and (p0 == p1).all()                 # every rank holds the same summed gradient and the same new parameters
    assert v0 == v1
    # single process: the two halves as two separate steps' gradients, averaged, one AdamW step
    from test_train_step_cpu import _train_inputs
    from torch_train_ops import TorchOps
    from unimm_b200.config import tiny_config
    from unimm_b200.train_step import TrainStep
    from unimm_b200.weights import random_state_dict
    cfg = tiny_config()
    sd = random_state_dict(cfg, seed=5, perturbed=True)
    _, _, batch = _train_inputs(4)
    ts = TrainStep(cfg, sd, TorchOps(), lr=1e-3, image_lr=1e-3, warmup_steps=0)
    va = ts.forward_backward(_slice(batch, 0, 2))
    ga = ts.params.g.clone()
    vb = ts.forward_backward(_slice(batch, 2, 4))
    gsum = ga + ts.params.g
    assert float((torch.from_numpy(g0) - gsum).abs().max()) < 1e-12 * float(gsum.abs().max())
    assert abs(v0["lm_loss"] - 0.5 * (va["lm_loss"] + vb["lm_loss"])) < 1e-9
    ts.params.g.copy_(0.5 * gsum)
    ts.optimizer_step()
    assert float((torch.from_numpy(p0) - ts.params.p).abs().max()) < 1e-12
