"""GPU: edge cases and error behaviour of the boundary — ragged and extreme batch shapes, the
capacity limits of an engine, and the loud failures the reference's silent or exception paths become."""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

if not torch.cuda.is_available():
    pytest.skip("needs a CUDA device", allow_module_level=True)

from unimm_b200 import synthetic as syn  # noqa: E402
from unimm_b200._lib import Config, UnimmError, lib  # noqa: E402
from unimm_b200.config import tiny_config  # noqa: E402
from unimm_b200.engine import Engine  # noqa: E402
from unimm_b200.packing import pack_units, units_from_rounds  # noqa: E402
from unimm_b200.weights import random_state_dict  # noqa: E402


@pytest.fixture(scope="module")
def tiny():
    cfg = tiny_config()
    sd = random_state_dict(cfg, seed=3, perturbed=True)
    eng = {p: Engine(cfg, sd, precision=p, max_sequences=64) for p in ("fp32", "fp16")}
    yield cfg, sd, eng
    for e in eng.values():
        e.close()


def _dense(eng, rounds, slots, feat, loc, mask, want=("seq_score", "nsp_scores")):
    tokens, segments, positions, labels, desc, _ = syn.stack_rounds(rounds)
    index = torch.tensor(np.concatenate([np.full(len(r.tokens), s, np.int32) for r, s in zip(rounds, slots)]))
    return eng.forward(tokens, segments, positions, desc, torch.from_numpy(feat), torch.from_numpy(loc), torch.from_numpy(mask),
                       feat_index=index, masked_lm_labels=labels, want=want)


@pytest.mark.parametrize("precision", ["fp32", "fp16"])
def test_ragged_and_extreme_units(tiny, precision):
    """One candidate in a round, a one-question context (round 1, short caption), the longest answers (last_len 8: 17 own rows),
    a context that fills the sequence to T = 256, units with different candidate counts — packed == dense for all of them."""
    _, _, engines = tiny
    eng = engines[precision]
    rng = np.random.RandomState(21)
    imgs = [syn.synth_image(rng) for _ in range(2)]
    feat, loc, mask = (np.stack([im[i] for im in imgs]) for i in range(3))
    mask[1, 30:] = 0                                                   # padded regions on the second image
    long_ctx = syn.synth_context(rng, 10, question_len=8)              # 240 context positions: T = 256 with the longest answer
    rounds = [
        syn.encode_round_gen(syn.synth_context(rng, 1, caption_len=1, question_len=1), syn.synth_answers(rng, 1)),
        syn.encode_round_gen(syn.synth_context(rng, 5), syn.synth_answers(rng, 13, len_range=(7, 7))),
        syn.encode_round_gen(long_ctx, syn.synth_answers(rng, 3, len_range=(7, 7))),
        syn.encode_round_gen(syn.synth_context(rng, 3), syn.synth_answers(rng, 40)),
    ]
    assert int(rounds[2].desc[0, 2] + rounds[2].desc[0, 3]) == 256     # L + last_len: the sequence is full
    slots = [0, 0, 1, 1]
    pb = pack_units(units_from_rounds(rounds, slots), feat, loc, mask)
    packed = eng.forward_packed(pb.to(eng.device), want=("seq_score", "nsp_scores"))
    dense = _dense(eng, rounds, slots, feat, loc, mask)
    d1 = (packed["seq_score"] - dense["seq_score"]).abs().max().item()
    d2 = (packed["nsp_scores"] - dense["nsp_scores"]).abs().max().item()
    print(f"[{precision}] extreme units: seq_score diff {d1:.3e}  nsp diff {d2:.3e}")
    tol = 5e-5 if precision == "fp32" else 2e-2
    assert torch.isfinite(packed["seq_score"]).all() and d1 < tol and d2 < tol
    # the same units packed for the scores only (no [CLS] / A_last rows; tail pruning inside the engine)
    lean = pack_units(units_from_rounds(rounds, slots), feat, loc, mask, scores_only=True)
    assert lean.n_text_rows == pb.n_text_rows - 3 * pb.n_cands + pb.n_units         # no [CLS] / A_last rows, one B_0 row per unit
    d3 = (eng.forward_packed(lean.to(eng.device), want=("seq_score",))["seq_score"] - dense["seq_score"]).abs().max().item()
    print(f"[{precision}] extreme units, scores-only packing: seq_score diff {d3:.3e}")
    assert d3 < tol


@pytest.mark.parametrize("precision", ["fp32", "fp16"])
def test_empty_answers_pack_to_a_single_row(tiny, precision):
    """Candidates whose answer is empty (last_len = 1: only the closing [SEP] is predicted) keep ONE row each in the scores-only
    layout (cand_halo = 0, no own-candidate keys besides the row itself); mixed with ordinary candidates in a second unit."""
    _, _, engines = tiny
    eng = engines[precision]
    rng = np.random.RandomState(33)
    feat, loc, mask = (a[None] for a in syn.synth_image(rng))
    empty = syn.encode_round_gen(syn.synth_context(rng, 2), [[] for _ in range(5)])
    mixed = syn.encode_round_gen(syn.synth_context(rng, 4), [[], [2000], [], [3000, 3001, 3002]])
    tol = 5e-5 if precision == "fp32" else 2e-2
    for rounds in ([empty], [empty, mixed]):
        slots = [0] * len(rounds)
        dense = _dense(eng, rounds, slots, feat, loc, mask, want=("seq_score",))["seq_score"]
        for scores_only in (False, True):
            pb = pack_units(units_from_rounds(rounds, slots), feat, loc, mask, scores_only=scores_only)
            if scores_only and len(rounds) == 1:
                assert pb.cand_halo == 0 and pb.n_text_rows == pb.n_shared_rows + 1      # five empty answers: the unit's B_0 row is all there is
            got = eng.forward_packed(pb.to(eng.device), want=("seq_score",))["seq_score"]
            d = (got - dense).abs().max().item()
            print(f"[{precision}] empty answers, {len(rounds)} unit(s), scores_only={scores_only}: seq_score diff {d:.3e}")
            assert torch.isfinite(got).all() and d < tol


def test_capacity_limits_fail_loudly(tiny):
    cfg, sd, engines = tiny
    eng = engines["fp16"]
    rng = np.random.RandomState(2)
    feat, loc, mask = (a[None] for a in syn.synth_image(rng))
    rnd = syn.encode_round_gen(syn.synth_context(rng, 4), syn.synth_answers(rng, 65))          # 65 > max_sequences = 64
    with pytest.raises(ValueError, match="exceeds max_sequences"):
        _dense(eng, [rnd], [0], feat, loc, mask)
    big = [syn.encode_round_gen(syn.synth_context(rng, 10), syn.synth_answers(rng, 100, len_range=(7, 7))) for _ in range(12)]
    pb = pack_units(units_from_rounds(big, [0] * 12), feat, loc, mask)                        # ~23k rows > 64 * 256
    assert pb.n_text_rows > 64 * 256
    with pytest.raises(UnimmError, match="exceeds the engine workspace"):
        eng.forward_packed(pb.to(eng.device), want=("seq_score",))
    # the engine is still usable afterwards
    ok = syn.encode_round_gen(syn.synth_context(rng, 2), syn.synth_answers(rng, 5))
    out = _dense(eng, [ok], [0], feat, loc, mask)
    assert torch.isfinite(out["seq_score"]).all()


def test_checkpoint_errors_name_the_key(tiny):
    cfg, sd, _ = tiny
    bad = dict(sd)
    key = next(k for k in bad if k.endswith("encoder.layer.0.output.dense.weight"))
    del bad[key]
    with pytest.raises(UnimmError, match="output.dense.weight"):
        Engine(cfg, bad, precision="fp16", max_sequences=4)
    wrong = dict(sd)
    wrong[key] = wrong[key][:-1]
    with pytest.raises(UnimmError, match="unexpected shape"):
        Engine(cfg, wrong, precision="fp16", max_sequences=4)


def test_c_abi_argument_checks():
    eng = C.c_void_p()
    assert lib.unimm_create(None, 0, 2, 4, C.byref(eng)) != 0 and b"null" in lib.unimm_last_error()
    cfg = Config()
    cfg.hidden_size, cfg.num_attention_heads = 700, 10                                         # not a supported width
    assert lib.unimm_create(C.byref(cfg), 0, 2, 4, C.byref(eng)) != 0 and b"hidden_size" in lib.unimm_last_error()
    assert lib.unimm_forward(None, None, None, None) != 0
    assert lib.unimm_destroy(None) == 0


def test_out_of_range_ids_do_not_fault(tiny):
    """The reference's nn.Embedding raises IndexError (and its assert forces a sync, vilbert_dialog.py:342); here ids are
    clamped on the device, the forward completes and stays finite."""
    _, _, engines = tiny
    eng = engines["fp16"]
    rng = np.random.RandomState(4)
    feat, loc, mask = (a[None] for a in syn.synth_image(rng))
    rnd = syn.encode_round_gen(syn.synth_context(rng, 2), syn.synth_answers(rng, 4))
    rnd.tokens[1, 3] = 10 ** 6
    rnd.positions[2, 5] = 5000
    out = _dense(eng, [rnd], [0], feat, loc, mask)
    torch.cuda.synchronize()
    assert torch.isfinite(out["seq_score"]).all()
