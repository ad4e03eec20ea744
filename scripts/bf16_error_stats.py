#!/usr/bin/env python
"""Error statistics (not only the maximum) of the 16-bit modes on the 300-candidate bench-shape fixtures, against the unmodified
reference's scores stored in tests/golden/sweep3x100_*.npz:  mode variants are selected by the engine's environment switches.
    python scripts/bf16_error_stats.py  > profiles/r02_v23_bf16_error_stats.txt        (GPU box)"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from unimm_b200 import synthetic as syn  # noqa: E402
from unimm_b200.config import DEFAULT_CONFIG_PATH, ViLBertConfig  # noqa: E402
from unimm_b200.engine import Engine  # noqa: E402
from unimm_b200.flat_packer import FlatPacker, ImageArrays  # noqa: E402
from unimm_b200.weights import random_state_dict  # noqa: E402

cfg = ViLBertConfig.from_json_file(DEFAULT_CONFIG_PATH)
VARIANTS = [("fp16", "fp16", {}), ("bf16 (default: mixed formats, fp16 residual stream)", "bf16", {}),
            ("bf16, fp32 residual stream", "bf16", {"UNIMM_RES16": "0"}), ("bf16 for every operand", "bf16", {"UNIMM_BF16_PURE": "1"})]
pk = FlatPacker()
print(f"{'variant':58s} {'fixture':22s}   max |err|    rms err    mean err   rank changes / 300")
for name in ("sweep3x100_default", "sweep3x100_perturbed"):
    g = dict(np.load(os.path.join(ROOT, "tests", "golden", name + ".npz")))
    sd = random_state_dict(cfg, int(g["weight_seed"]), perturbed=bool(g["perturbed"]))
    (f, l, m), rs = syn.synth_dialog_rounds(int(g["image_id"]), rounds=tuple(int(r) for r in g["round_ids"]))
    view = pk.pack([ImageArrays.from_rounds(rs, f, l, m)], scores_only=True, share_first_mask=True, verify_shared=True)
    for label, prec, env in VARIANTS:
        for k in ("UNIMM_RES16", "UNIMM_BF16_PURE"):
            os.environ.pop(k, None)
        os.environ.update(env)
        eng = Engine(cfg, sd, precision=prec, max_sequences=8 * 52)
        out = torch.zeros(300).pin_memory()
        eng.score_packed_host(view, out)
        eng.close()
        e = out.numpy().reshape(3, 100) - g["seq_score"]
        order = np.argsort(-out.numpy().reshape(3, 100), axis=1, kind="stable")
        ranks = np.empty_like(order)
        np.put_along_axis(ranks, order, np.arange(1, 101)[None].repeat(3, 0), axis=1)
        print(f"{label:58s} {name:22s}   {np.abs(e).max():.3e}  {np.sqrt((e ** 2).mean()):.3e}  {e.mean():+.3e}   {int((ranks != g['ranks']).sum())}")
pk.close()
