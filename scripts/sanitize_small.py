"""Small end-to-end calls of every round-2 kernel path for compute-sanitizer (memcheck / racecheck): tiny 2-layer config, dense and
packed layouts in the three precision modes, LM-head backward on a small vocabulary slice.
    compute-sanitizer --tool memcheck python scripts/sanitize_small.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from unimm_b200 import synthetic as syn  # noqa: E402
from unimm_b200.config import tiny_config  # noqa: E402
from unimm_b200.engine import Engine  # noqa: E402
from unimm_b200.flat_packer import FlatPacker, ImageArrays  # noqa: E402
from unimm_b200.lm_head_grad import lm_head_backward  # noqa: E402
from unimm_b200.weights import random_state_dict  # noqa: E402

cfg = tiny_config()
sd = random_state_dict(cfg, seed=3, perturbed=True)
rng = np.random.RandomState(0)
(feat, loc, mask), rounds = syn.synth_dialog_rounds(5, rounds=(1, 4), n_candidates=6)
tokens, segments, positions, labels, desc, index = syn.stack_rounds(rounds)
index = torch.zeros_like(index)
im = ImageArrays.from_rounds(rounds, feat, loc, mask)
pk = FlatPacker()
for prec in ("fp32", "fp16", "bf16"):
    eng = Engine(cfg, sd, precision=prec, max_sequences=16)
    o = eng.forward(tokens, segments, positions, desc, torch.from_numpy(feat)[None], torch.from_numpy(loc)[None], torch.from_numpy(mask)[None],
                    feat_index=index, masked_lm_labels=labels, want=("seq_score", "nsp_scores"))
    eng.check_ids()
    dense = o["seq_score"].cpu()
    for so in (True, False):
        v = pk.pack([im], scores_only=so)
        out, nsp = torch.zeros(v.n_cands).pin_memory(), torch.zeros(v.n_cands, 2).pin_memory()
        eng.submit_packed_host(v, 1, out, None if so else nsp)
        eng.wait_packed(1)
        print(prec, "scores_only" if so else "with cls", "packed vs dense", float((out - dense).abs().max()))
    eng.close()
g = torch.Generator().manual_seed(0)
n, V, K = 70, 1000, 768
h, E = torch.randn(n, K, generator=g), 0.05 * torch.randn(V, K, generator=g)
w = torch.ones(n); w[1::3] = -1
out = lm_head_backward(h.cuda(), E.cuda(), torch.zeros(V), torch.randint(0, V, (n,), generator=g), w, grad_scale=1.0 / n)
torch.cuda.synchronize()
print("lm head backward", float(out["dH"].abs().max()), float(out["dE"].abs().max()))
print("sanitize_small done")
