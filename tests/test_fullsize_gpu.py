"""GPU: the hot path at BASELINE.json's bench size (one step = 8 images x 10 rounds x 100 candidates = 8000 candidates,
~98k packed text rows), checked through size-independent properties — the oracle cannot run this size in seconds.

  * prefix-shared (packed) scoring == the dense per-sequence forward the reference computes (same weights, same kernels
    for the projections; different row layout, attention kernels and GEMM tile shapes: only 16-bit rounding differs)
  * permuting the candidates of every round permutes the scores
  * a duplicated candidate gets the score of its twin

This is also the only place the tests drive the kernels that exist for tall problems: the CTA-pair (multicast) GEMM,
the tcgen05 candidate / cross attention over thousands of tiles, the LayerNorm cluster kernel over many waves.
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

if not torch.cuda.is_available():
    pytest.skip("needs a CUDA device", allow_module_level=True)

from unimm_b200 import synthetic as syn  # noqa: E402
from unimm_b200.engine import Engine  # noqa: E402
from unimm_b200.packing import pack_units, units_from_rounds  # noqa: E402
from unimm_b200.weights import random_state_dict  # noqa: E402

N_IMAGES = 8
TOL = 2e-2          # the north star's 16-bit bound on a sequence log-likelihood


@pytest.fixture(scope="module")
def step():
    rounds, slots, feats, locs, masks = [], [], [], [], []
    for i in range(N_IMAGES):
        (feat, loc, mask), rs = syn.synth_dialog_rounds(7000 + i)
        rounds += rs
        slots += [i] * len(rs)
        feats.append(feat), locs.append(loc), masks.append(mask)
    return rounds, slots, np.stack(feats), np.stack(locs), np.stack(masks)


@pytest.fixture(scope="module")
def engine(full_cfg, step):
    rounds, slots, feat, loc, mask = step
    pb = pack_units(units_from_rounds(rounds, slots), feat, loc, mask)
    cap = max(-(-pb.n_text_rows // 256), pb.n_units, 250) + 1
    eng = Engine(full_cfg, random_state_dict(full_cfg, 0), precision="fp16", max_sequences=cap)
    yield eng
    eng.close()


def packed_scores(eng, rounds, slots, feat, loc, mask, scores_only=False):
    pb = pack_units(units_from_rounds(rounds, slots), feat, loc, mask, scores_only=scores_only)
    out = eng.forward_packed(pb.to(eng.device), want=("seq_score",))["seq_score"]
    torch.cuda.synchronize()
    return out.cpu().numpy(), pb


def test_packed_equals_dense_at_bench_size(engine, step):
    rounds, slots, feat, loc, mask = step
    packed, pb = packed_scores(engine, rounds, slots, feat, loc, mask)
    assert packed.shape == (N_IMAGES * 1000,) and np.isfinite(packed).all()
    assert pb.n_text_rows > 90000                                     # tall enough for the CTA-pair GEMM and multi-wave kernels
    tokens, segments, positions, labels, desc, _ = syn.stack_rounds(rounds)
    index = torch.tensor(np.concatenate([np.full(len(r.tokens), s, np.int32) for r, s in zip(rounds, slots)]))
    dense = np.zeros_like(packed)
    f, l, m = torch.from_numpy(feat), torch.from_numpy(loc), torch.from_numpy(mask)
    for s in range(0, len(dense), 250):
        e = s + 250
        o = engine.forward(tokens[s:e], segments[s:e], positions[s:e], desc[s:e], f, l, m, feat_index=index[s:e],
                           masked_lm_labels=labels[s:e], want=("seq_score",))
        dense[s:e] = o["seq_score"].cpu().numpy()
    diff = np.abs(packed - dense)
    # ranking agreement per round: the top candidate of the dense path is the top (or within the tolerance of it) when packed
    p, d = packed.reshape(-1, 100), dense.reshape(-1, 100)
    top_gap = p.max(1) - p[np.arange(p.shape[0]), d.argmax(1)]
    print(f"packed vs dense over {len(dense)} candidates ({pb.n_text_rows} packed rows): max |diff| {diff.max():.3e}, mean {diff.mean():.3e}, "
          f"top-1 agreement {(p.argmax(1) == d.argmax(1)).mean():.3f}")
    assert diff.max() < TOL
    assert (top_gap < TOL).all()
    # scores-only packing (what bench.py and the sweep driver run): 2 rows fewer per candidate, same scores
    lean, pl = packed_scores(engine, rounds, slots, feat, loc, mask, scores_only=True)
    assert pl.n_text_rows == pb.n_text_rows - 3 * pb.n_cands + pb.n_units      # no [CLS] / A_last rows, one B_0 row per unit
    d_lean = np.abs(lean - dense)
    print(f"scores-only packing ({pl.n_text_rows} rows) vs dense: max |diff| {d_lean.max():.3e}; vs full packing {np.abs(lean - packed).max():.3e}")
    assert d_lean.max() < TOL and np.abs(lean - packed).max() < 5e-3


def test_candidate_permutation_and_duplicates(engine, step):
    rounds, slots, feat, loc, mask = step
    base, _ = packed_scores(engine, rounds, slots, feat, loc, mask)
    rng = np.random.RandomState(3)
    perm_rounds, perms = [], []
    for r in rounds:
        p = rng.permutation(len(r.tokens))
        p[1] = p[0]                                                   # candidate 1 becomes a copy of candidate 0 (after permutation)
        perms.append(p)
        perm_rounds.append(syn.Round(r.tokens[p], r.segments[p], r.positions[p], r.labels[p], r.desc[p]))
    got, _ = packed_scores(engine, perm_rounds, slots, feat, loc, mask)
    want = np.concatenate([base[100 * u:100 * (u + 1)][p] for u, p in enumerate(perms)])
    d_perm = np.abs(got - want).max()
    g = got.reshape(-1, 100)
    d_dup = np.abs(g[:, 0] - g[:, 1]).max()
    print(f"permutation: max |diff| {d_perm:.3e}; duplicate candidates: max |diff| {d_dup:.3e}")
    # rows move to other tiles / warps: only the accumulation order inside the 16-bit kernels changes
    assert d_perm < 5e-3
    assert d_dup < 5e-3
