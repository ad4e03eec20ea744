"""GPU parity of the BENCH CONFIGURATION against the unmodified reference: the packed, scores-only, shared-B_0, LM-dedup path
through the C++ packer and unimm_score_packed_host, on a full 8-image step that contains three rounds (1, 5, 10 — contexts of
different lengths) of one image whose 300 sequence log-likelihoods the reference computed (tests/golden/sweep3x100_*.npz,
made by tests/golden/make_golden.py as val_lm.py:104-137 does: chunks of 25, full logits, cross_entropy).  Plus the sequences
the reference truncates at max_seq_len (tests/golden/gen10_truncated.npz, utils/data_utils.py:205-209, :237-244)."""
import numpy as np
import pytest
import torch

from conftest import golden_state_dict, load_golden

pytestmark = pytest.mark.gpu

if not torch.cuda.is_available():
    pytest.skip("needs a CUDA device", allow_module_level=True)

from unimm_b200 import synthetic as syn  # noqa: E402
from unimm_b200.descriptors import descriptors_from_masks  # noqa: E402
from unimm_b200.engine import Engine  # noqa: E402
from unimm_b200.flat_packer import FlatPacker, ImageArrays  # noqa: E402

TOL = {"fp32": 1e-4, "fp16": 2e-2, "bf16": 2e-2}            # BASELINE.json north_star — not widened for any mode
# Measured on these 300-candidate fixtures (maximum over the 100 candidates of a round): fp32 3e-5, fp16 5e-3, bf16 1.2e-2 .. 1.9e-2
# (1.2e-2 .. 1.5e-2 with the fp32 residual stream, UNIMM_RES16=0: same mean error, 11 % slower).
# The bf16 mode writes LayerNorm outputs — bounded by construction — and the weights of the projections that read them as fp16, and
# everything of unbounded range (Q / K / V, attention context, GELU outputs, image features, their weights) as bf16; the residual
# stream is the fp16 LayerNorm output, as in the fp16 mode (engine.cu: mix16, res16).  With bf16 for EVERY operand (UNIMM_BF16_PURE=1, round 2's first version) the same cases
# measure 2.0e-2 .. 2.3e-2, at / above the bound; tests/bf16_rounding_study.py is the CPU study that located the error.
_ENG = {}


def engine(cfg, seed, perturbed, precision, max_sequences):
    key = (int(seed), bool(perturbed), precision, max_sequences)
    if key not in _ENG:
        for e in _ENG.values():
            e.close()
        _ENG.clear()
        torch.cuda.empty_cache()
        _ENG[key] = Engine(cfg, golden_state_dict(cfg, seed, perturbed), precision=precision, max_sequences=max_sequences)
    return _ENG[key]


@pytest.mark.parametrize("precision", ["fp32", "fp16", "bf16"])
@pytest.mark.parametrize("name", ["sweep3x100_perturbed", "sweep3x100_default"])
def test_bench_step_reproduces_reference_scores(full_cfg, name, precision):
    from oracle import visdial_metrics as om
    import os

    from conftest import GOLDEN_DIR
    g = dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz")))
    image_id, round_ids = int(g["image_id"]), tuple(int(r) for r in g["round_ids"])
    (f, l, m), rs = syn.synth_dialog_rounds(image_id, rounds=round_ids)
    # the fixture pins the generator: these are the very arrays the reference scored
    assert np.array_equal(np.concatenate([r.tokens for r in rs]), g["tokens"]) and np.array_equal(np.concatenate([r.labels for r in rs]), g["labels"])
    assert np.array_equal(np.concatenate([r.desc for r in rs]), g["desc"])
    images = []
    for i in (3, 11):                                          # full 10-round images before and after, as in a bench step
        (f2, l2, m2), r2 = syn.synth_dialog_rounds(i)
        images.append(ImageArrays.from_rounds(r2, f2, l2, m2))
    images.insert(1, ImageArrays.from_rounds(rs, f, l, m))
    for i in (20, 21, 22, 23, 24):
        (f2, l2, m2), r2 = syn.synth_dialog_rounds(i)
        images.append(ImageArrays.from_rounds(r2, f2, l2, m2))
    pk = FlatPacker()
    view = pk.pack(images, scores_only=True, share_first_mask=True, verify_shared=True)
    assert view.n_units == 73 and view.n_cands == 7300 and view.struct.n_lm_unique > 0 and view.struct.n_images == 8
    eng = engine(full_cfg, g["weight_seed"], g["perturbed"], precision, 8 * 52)
    out = torch.zeros(view.n_cands).pin_memory()
    eng.score_packed_host(view, out)
    mine = out[1000:1300].view(3, 100)
    err = (mine.numpy() - g["seq_score"])
    flips = int((om.scores_to_ranks(mine.view(1, 3, 100)).view(3, 100).numpy() != g["ranks"]).sum())
    print(f"[{precision}] {name}: bench-shape step, rounds {round_ids}: max |seq_score err| per round {np.abs(err).max(1)}, rank changes {flips}/300")
    assert np.abs(err).max() < TOL[precision]
    if precision == "fp32":
        assert flips == 0
    # the same units alone (no neighbours in the batch) give the same numbers: units are independent
    v2 = pk.pack([images[1]], scores_only=True)
    out2 = torch.zeros(300).pin_memory()
    eng.score_packed_host(v2, out2)
    assert (out2 - out[1000:1300]).abs().max().item() < (1e-5 if precision == "fp32" else 6e-3)
    pk.close()


@pytest.mark.parametrize("switch", ["UNIMM_LM_HP", "UNIMM_BF16_PURE"])
def test_bf16_switches(full_cfg, switch, monkeypatch):
    """UNIMM_LM_HP=1: the LM head of the bf16 mode at the fp32-class precision (split3 over fp16 hi | lo planes, fed from the fp32
    residual stream) — inside the bound, and not the same numbers as the default.  UNIMM_BF16_PURE=1: bf16 for every operand — the
    variant the default replaced; it must still run and stay within 3e-2 (it measures 2.0e-2 .. 2.3e-2: NOT a north-star claim)."""
    import os

    from conftest import GOLDEN_DIR
    g = dict(np.load(os.path.join(GOLDEN_DIR, "sweep3x100_default.npz")))
    (f, l, m), rs = syn.synth_dialog_rounds(int(g["image_id"]), rounds=tuple(int(r) for r in g["round_ids"]))
    pk = FlatPacker()
    view = pk.pack([ImageArrays.from_rounds(rs, f, l, m)], scores_only=True, share_first_mask=True, verify_shared=True)
    base = torch.zeros(300).pin_memory()
    engine(full_cfg, g["weight_seed"], g["perturbed"], "bf16", 8 * 52).score_packed_host(view, base)
    monkeypatch.setenv(switch, "1")
    for e in _ENG.values():
        e.close()
    _ENG.clear()                                               # the switches are read when the engine is created
    out = torch.zeros(300).pin_memory()
    engine(full_cfg, g["weight_seed"], g["perturbed"], "bf16", 8 * 52).score_packed_host(view, out)
    for e in _ENG.values():
        e.close()
    _ENG.clear()
    err = np.abs(out.numpy().reshape(3, 100) - g["seq_score"]).max()
    print(f"[bf16, {switch}=1] max |seq_score err| {err:.3e}; max |difference from the default| {(out - base).abs().max().item():.3e}")
    assert err < (TOL["bf16"] if switch == "UNIMM_LM_HP" else 3e-2)
    assert (out - base).abs().max().item() > 1e-4
    pk.close()


@pytest.mark.parametrize("precision", ["fp32", "fp16", "bf16"])
def test_truncated_sequences_match_reference(full_cfg, precision):
    g, b = load_golden("gen10_truncated")
    desc = descriptors_from_masks(b["txt_attention_mask"], b["co_attention_mask"])
    eng = engine(full_cfg, g["weight_seed"], g["perturbed"], precision, 16)
    # dense layout (what VisualDialogEncoder.forward runs): descriptors with L + last_len > S
    o = eng.forward(b["tokens"], b["segments"], b["positions"], desc, b["image_feat"], b["image_loc"], b["image_mask"],
                    masked_lm_labels=b["mask"], want=("seq_score", "nsp_scores", "token_logp"))
    seq = o["seq_score"].cpu().numpy()
    rows = g["token_rows"]
    tl = o["token_logp"].cpu().numpy()[rows[:, 0], rows[:, 1]]
    e1, e2, e3 = np.abs(seq - g["seq_score"]).max(), np.abs(tl - g["token_logp"]).max(), np.abs(o["nsp_scores"].cpu().numpy() - g["nsp_scores"]).max()
    print(f"[{precision}] truncated, dense layout: seq_score err {e1:.3e} token_logp err {e2:.3e} nsp err {e3:.3e}")
    assert max(e1, e2, e3) < TOL[precision]
    assert seq[5] == 0 and seq[6] == 0                         # no masked-copy position inside S: an empty sum, as val_lm.py:131-136
    # packed layout, descriptors derived by the packer from the position ids
    tok, sg, ps, lab = (b[k].numpy() for k in ("tokens", "segments", "positions", "mask"))
    pk = FlatPacker()
    for scores_only in (True, False):
        v = pk.pack([ImageArrays(tok, sg, ps, lab, g["image_feat"], g["image_loc"], g["image_mask"], units=[(0, 10)])], scores_only=scores_only)
        out, nsp = torch.zeros(10).pin_memory(), torch.zeros(10, 2).pin_memory()
        eng.score_packed_host(v, out, None if scores_only else nsp)
        err = np.abs(out.numpy() - g["seq_score"]).max()
        print(f"[{precision}] truncated, packed layout (scores_only={scores_only}): seq_score err {err:.3e}")
        assert err < TOL[precision]
        if not scores_only:
            assert np.abs(nsp.numpy() - g["nsp_scores"]).max() < TOL[precision]
    pk.close()


def test_out_of_range_ids_are_an_error(full_cfg):
    """The reference's nn.Embedding raises on an id outside its table; the engine must not return plausible scores."""
    from unimm_b200._lib import UnimmError
    g, b = load_golden("gen8_default")
    eng = engine(full_cfg, g["weight_seed"], g["perturbed"], "fp16", 16)
    tok, sg, ps, lab = (b[k].numpy().copy() for k in ("tokens", "segments", "positions", "mask"))
    pk = FlatPacker()
    im = lambda t, s_, p_: ImageArrays(t, s_, p_, lab, g["image_feat"], g["image_loc"], g["image_mask"], units=[(0, 8)])
    out = torch.zeros(8).pin_memory()
    eng.score_packed_host(pk.pack([im(tok, sg, ps)]), out)                      # clean batch: fine
    bad = tok.copy()
    bad[:, 5] = 30522
    with pytest.raises(UnimmError, match="outside its embedding table"):
        eng.score_packed_host(pk.pack([im(bad, sg, ps)]), out)
    bad = sg.copy()
    bad[:, 3] = 12
    with pytest.raises(UnimmError, match="outside its embedding table"):
        eng.score_packed_host(pk.pack([im(tok, bad, ps)]), out)
    eng.score_packed_host(pk.pack([im(tok, sg, ps)]), out)                      # and the flag does not stick
    np.testing.assert_allclose(out.numpy(), g["seq_score"], atol=2e-2)
    pk.close()


@pytest.mark.parametrize("precision", ["fp32", "fp16", "bf16"])
def test_config4_nsp_ranking_of_100_candidates(full_cfg, precision):
    """BASELINE config 4 at its stated size (val.py:125-131): 100 options under the discriminative masks, ranked by the NSP
    probability softmax(seq_relationship_score)[:, 0]."""
    g, b = load_golden("dis100_default")
    desc = descriptors_from_masks(b["txt_attention_mask"], b["co_attention_mask"])
    eng = engine(full_cfg, g["weight_seed"], g["perturbed"], precision, 128)
    o = eng.forward(b["tokens"], b["segments"], b["positions"], desc, torch.from_numpy(g["image_feat"])[None], torch.from_numpy(g["image_loc"])[None],
                    torch.from_numpy(g["image_mask"])[None], feat_index=torch.zeros(100, dtype=torch.int32), want=("nsp_scores",))
    eng.check_ids()
    nsp = o["nsp_scores"].cpu()
    p0 = torch.softmax(nsp, 1)[:, 0].numpy()
    e1, e2 = np.abs(nsp.numpy() - g["nsp_scores"]).max(), np.abs(p0 - g["nsp_prob0"]).max()
    print(f"[{precision}] config 4, 100 options: nsp logit err {e1:.3e}, P(answer) err {e2:.3e}")
    assert e1 < TOL[precision] and e2 < TOL[precision]


@pytest.mark.parametrize("precision", ["fp32", "fp16", "bf16"])
def test_config3_train_step_forward_at_size(full_cfg, precision):
    """BASELINE config 3 at its stated size (train.py:53-92, :445): 240 sequences = 40 images x (1 positive + 5 negatives), mixed
    generative / discriminative masks, 15 % masking, unlikelihood on the negatives: the three losses of the reference."""
    from conftest import load_golden_multi_image
    g, b, im = load_golden_multi_image("train240_perturbed")
    desc = descriptors_from_masks(b["txt_attention_mask"], b["co_attention_mask"])
    assert 60 < int((desc[:, 0] == 1).sum()) < 180                  # both mask families in one batch
    eng = engine(full_cfg, g["weight_seed"], g["perturbed"], precision, 240)
    idx = b["seq_image"]
    o = eng.forward(b["tokens"], b["segments"], b["positions"], desc, im["image_feat"], im["image_loc"], im["image_mask"],
                    feat_index=idx.to(torch.int32), masked_lm_labels=b["mask"], lm_weight=b["weights"],
                    next_sentence_label=torch.from_numpy(g["next_sentence_label"]), image_label=im["image_label"][idx].contiguous(),
                    image_target=im["image_target"][idx].contiguous(), nsp_weight=torch.from_numpy(g["nsp_weight"]), want=("losses", "nsp_scores"))
    eng.check_ids()
    lm, img, nsp = o["losses"][:3].cpu().numpy()
    print(f"[{precision}] config 3, B = 240: lm {lm:.6f}/{g['lm_loss'].item():.6f} img {img:.6f}/{g['img_loss'].item():.6f} "
          f"nsp {nsp:.6f}/{g['nsp_loss'].item():.6f}; nsp logits err {np.abs(o['nsp_scores'].cpu().numpy() - g['nsp_scores']).max():.3e}")
    tol = TOL[precision]
    assert abs(lm - g["lm_loss"].item()) < tol and abs(img - g["img_loss"].item()) < tol and abs(nsp - g["nsp_loss"].item()) < tol
    np.testing.assert_allclose(o["nsp_scores"].cpu().numpy(), g["nsp_scores"], atol=tol, rtol=0)


@pytest.mark.parametrize("precision", ["fp32", "fp16"])
@pytest.mark.parametrize("name", ["ft100gen_perturbed", "ft100dis_perturbed"])
def test_config5_dense_annotation_step_at_size(full_cfg, name, precision):
    """BASELINE config 5 at its stated size (dense_annotation_finetuning.py:253-296): the 100 options of one annotated round,
    relevance-weighted L / UL loss, unweighted NSP CE, and the NeuralNDCG objective on the NSP probabilities."""
    from unimm_b200.rank_loss import neural_ndcg_loss
    g, b = load_golden(name)
    desc = descriptors_from_masks(b["txt_attention_mask"], b["co_attention_mask"])
    eng = engine(full_cfg, g["weight_seed"], g["perturbed"], precision, 128)
    n = 100
    o = eng.forward(b["tokens"], b["segments"], b["positions"], desc, torch.from_numpy(g["image_feat"])[None], torch.from_numpy(g["image_loc"])[None],
                    torch.from_numpy(g["image_mask"])[None], feat_index=torch.zeros(n, dtype=torch.int32), masked_lm_labels=b["mask"],
                    lm_weight=b["weights"], next_sentence_label=torch.from_numpy(g["next_sentence_label"]),
                    image_label=torch.from_numpy(g["image_label"]).unsqueeze(0).expand(n, -1).contiguous(),
                    image_target=torch.from_numpy(g["image_target"]).unsqueeze(0).expand(n, -1, -1).contiguous(), nsp_weight=None,
                    want=("losses", "nsp_scores"))
    lm, img, nsp = o["losses"][:3].cpu().numpy()
    probs = torch.softmax(o["nsp_scores"], -1)[:, 0].view(1, n)
    ndcg = float(neural_ndcg_loss(probs, torch.from_numpy(g["relevance"]).view(1, n).to(probs.device)))
    total = ndcg + lm + nsp
    print(f"[{precision}] {name}: lm {lm:.6f}/{g['lm_loss'].item():.6f} nsp {nsp:.6f}/{g['nsp_ce_unweighted'].item():.6f} "
          f"neuralNDCG {ndcg:.6f}/{float(g['neural_ndcg_loss']):.6f} total {total:.6f}/{float(g['total_loss']):.6f}")
    tol = TOL[precision]
    assert abs(lm - g["lm_loss"].item()) < tol and abs(img - g["img_loss"].item()) < tol
    assert abs(nsp - g["nsp_ce_unweighted"].item()) < tol            # nsp_weight None: the engine's NSP loss IS the unweighted CE
    assert abs(ndcg - float(g["neural_ndcg_loss"])) < (2e-4 if precision == "fp32" else 2e-2)
    assert abs(total - float(g["total_loss"])) < 3 * tol


def test_two_steps_in_flight_equal_blocking_calls(full_cfg):
    """unimm_submit_packed_host / unimm_wait_packed (step i + 1 queued in the second staging slot behind step i) return exactly
    what the blocking unimm_score_packed_host returns for each step, in either slot, across repeated use."""
    g, _ = load_golden("gen8_default")
    eng = engine(full_cfg, g["weight_seed"], g["perturbed"], "fp16", 64)
    steps = []
    for i in (1, 2, 3):
        (f, l, m), rs = syn.synth_dialog_rounds(60 + i, rounds=(1 + i, 9), n_candidates=20 + 5 * i)
        steps.append([ImageArrays.from_rounds(rs, f, l, m)])
    packers = [FlatPacker() for _ in range(3)]
    want = []
    for k, st in enumerate(steps):
        v = packers[0].pack(st)
        out = torch.zeros(v.n_cands).pin_memory()
        eng.score_packed_host(v, out)
        want.append(out.clone())
    outs = [torch.zeros(4096).pin_memory() for _ in range(2)]
    got, pending = [], None
    for rep in range(2):
        for k, st in enumerate(steps):
            i = rep * len(steps) + k
            v = packers[i % 3].pack(st)
            eng.submit_packed_host(v, i & 1, outs[i & 1])
            if pending is not None:
                eng.wait_packed(pending[0])
                got.append(outs[pending[0]][:pending[1]].clone())
            pending = (i & 1, v.n_cands)
    eng.wait_packed(pending[0])
    got.append(outs[pending[0]][:pending[1]].clone())
    for i, o in enumerate(got):
        assert torch.equal(o, want[i % len(steps)]), i
    for p in packers:
        p.close()
