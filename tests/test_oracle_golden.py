"""CPU: the oracle restatement against fixtures produced by the unmodified reference (tests/golden/)."""
import os

import numpy as np
import pytest
import torch

from conftest import golden_state_dict, load_golden
from oracle import encode_inputs as enc
from oracle import vilbert_oracle as vo
from oracle import visdial_metrics as om

# the oracle calls the same ATen CPU ops as the reference in the same order; only GEMM blocking may differ
TOL = 2e-5


def _run(cfg, g, batch, **kw):
    sd = golden_state_dict(cfg, g["weight_seed"], g["perturbed"])
    with torch.no_grad():
        return vo.forward(sd, cfg, batch["tokens"], batch["image_feat"], batch["image_loc"], batch["segments"],
                          batch["positions"], batch["txt_attention_mask"], batch["image_mask"],
                          batch["co_attention_mask"], masked_lm_labels=batch["mask"], lm_weight=batch["weights"], **kw)


@pytest.mark.parametrize("name", ["gen8_perturbed", "gen8_default"])
def test_generative_scores_match_reference(full_cfg, name):
    g, batch = load_golden(name)
    taps = {}
    out = _run(full_cfg, g, batch, taps=taps)
    assert np.array_equal(out["token_rows"].numpy(), g["token_rows"])
    np.testing.assert_allclose(out["token_logp"].numpy(), g["token_logp"], atol=TOL, rtol=0)
    np.testing.assert_allclose(out["token_ul"].numpy(), g["token_ul"], atol=TOL, rtol=0)
    np.testing.assert_allclose(out["seq_score"].numpy(), g["seq_score"], atol=5 * TOL, rtol=0)
    np.testing.assert_allclose(out["nsp_scores"].numpy(), g["nsp_scores"], atol=TOL, rtol=0)
    if "tap_txt_rows" in g:   # intermediate activations of sequence 0 (bisecting aid, pins every layer)
        trows, irows = g["tap_txt_rows"], g["tap_img_rows"]
        valid = trows < int(batch["txt_attention_mask"][0, 0].sum())      # pad rows are garbage-by-design
        for k in [k for k in g if k.startswith("tap.") and k not in ("tap_txt_rows", "tap_img_rows")]:
            key = k[4:]
            if key.endswith(".img") or key.startswith("v"):
                mine = taps[key if key.endswith(".img") else key + ".img"][0][irows]
                ref = g[k]
            else:
                mine = taps[key if key.endswith(".txt") else key + ".txt"][0][trows][valid]
                ref = g[k][valid]
            np.testing.assert_allclose(mine.numpy(), ref, atol=1e-4, rtol=0, err_msg=k)


def test_full_logits_path_equals_gathered_path(full_cfg):
    g, batch = load_golden("gen8_perturbed")
    b2 = {k: v[:2] for k, v in batch.items()}
    sd = golden_state_dict(full_cfg, g["weight_seed"], g["perturbed"])
    s_full, _, t_full = vo.score_candidates(sd, full_cfg, b2, chunk=2, full_logits=True)
    np.testing.assert_allclose(s_full.numpy(), g["seq_score"][:2], atol=5 * TOL, rtol=0)


def test_discriminative_nsp_matches_reference(full_cfg):
    g, batch = load_golden("dis8_perturbed")
    out = _run(full_cfg, g, batch)
    np.testing.assert_allclose(out["nsp_scores"].numpy(), g["nsp_scores"], atol=TOL, rtol=0)
    np.testing.assert_allclose(torch.softmax(out["nsp_scores"], 1)[:, 0].numpy(), g["nsp_prob0"], atol=TOL, rtol=0)
    np.testing.assert_allclose(out["token_logp"].numpy(), g["token_logp"], atol=TOL, rtol=0)


def test_training_losses_match_reference(full_cfg):
    g, batch = load_golden("train6_perturbed")
    n = batch["tokens"].shape[0]
    out = _run(full_cfg, g, batch,
               next_sentence_label=torch.from_numpy(g["next_sentence_label"]),
               image_label=torch.from_numpy(g["image_label"]).unsqueeze(0).expand(n, -1),
               image_target=torch.from_numpy(g["image_target"]).unsqueeze(0).expand(n, -1, -1),
               nsp_weight=torch.from_numpy(g["nsp_weight"]))
    np.testing.assert_allclose(out["lm_loss"].item(), g["lm_loss"].item(), atol=TOL, rtol=0)
    np.testing.assert_allclose(out["nsp_loss"].item(), g["nsp_loss"].item(), atol=TOL, rtol=0)
    np.testing.assert_allclose(out["img_loss"].item(), g["img_loss"].item(), atol=TOL, rtol=0)


@pytest.mark.parametrize("name", ["ft8gen_perturbed", "ft8dis_perturbed"])
def test_dense_annotation_losses_match_reference(full_cfg, name):
    """BASELINE config 5: one mask mode for all options, relevance as (integer-truncated) token weight, nsp_weight None."""
    g, batch = load_golden(name)
    n = batch["tokens"].shape[0]
    out = _run(full_cfg, g, batch,
               next_sentence_label=torch.from_numpy(g["next_sentence_label"]),
               image_label=torch.from_numpy(g["image_label"]).unsqueeze(0).expand(n, -1),
               image_target=torch.from_numpy(g["image_target"]).unsqueeze(0).expand(n, -1, -1),
               nsp_weight=None)
    np.testing.assert_allclose(out["lm_loss"].item(), g["lm_loss"].item(), atol=TOL, rtol=0)
    np.testing.assert_allclose(out["nsp_loss"].item(), g["nsp_loss"].item(), atol=TOL, rtol=0)
    np.testing.assert_allclose(out["img_loss"].item(), g["img_loss"].item(), atol=TOL, rtol=0)
    np.testing.assert_allclose(out["nsp_scores"].numpy(), g["nsp_scores"], atol=TOL, rtol=0)


def test_encoders_reproduce_reference_tensors():
    """oracle.encode_inputs regenerates the committed reference-made inputs from the same RNG stream."""
    g, batch = load_golden("gen8_default")
    rng = np.random.RandomState(1234)
    context, answers = enc.synth_round(rng, n_candidates=100)
    feats, loc, image_mask = enc.synth_image(rng)
    for j, n in enumerate([1, 2, 3, 4, 5, 6, 7, 4]):
        answers[j] = rng.randint(1000, 30522, size=n).tolist()
    r = np.random.RandomState(7)
    b = enc.build_batch(context, answers[:8], feats, loc, image_mask, mode="gen", rng=r)
    for k in ("tokens", "segments", "positions", "sep_indices", "mask", "weights"):
        assert torch.equal(b[k], batch[k]), k
    assert torch.equal(b["txt_attention_mask"].bool(), batch["txt_attention_mask"].bool())
    assert torch.equal(b["co_attention_mask"], batch["co_attention_mask"])
    np.testing.assert_array_equal(b["image_feat"][0].numpy(), g["image_feat"])


def test_rank_metrics_on_golden_scores():
    g, _ = load_golden("gen100_default")
    score = torch.from_numpy(g["seq_score"])
    assert np.array_equal(om.scores_to_ranks(score.view(1, 1, 100)).view(100).numpy(), g["ranks"])
    mine = {**om.sparse_metrics(score.view(1, 1, 100), torch.zeros(1, 1, dtype=torch.long)),
            "ndcg": om.ndcg(score.view(1, 100), torch.from_numpy(g["relevance"]))}
    for k, v in zip(g["metric_names"], g["metric_values"]):
        assert abs(mine[str(k)] - v) < 1e-6, k


def test_rank_loss_oracle_matches_reference():
    """NeuralNDCG-transposed and the NSP ensemble normalisation against values from the unmodified reference."""
    from oracle import rank_loss as orl
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "rankloss.npz"))
    for p, y, want in zip(z["y_pred"], z["y_true"], z["loss"]):
        got = orl.neural_ndcg_transposed(p, y)[0]
        assert abs(float(got) - float(want)) < 1e-6
    ens = orl.ensemble_normalise(z["ens_probs"].reshape(5, 40, 100)).reshape(4, 10, 100)
    np.testing.assert_allclose(ens, z["ens_out"], atol=1e-6, rtol=0)


def test_truncated_sequences_match_reference(full_cfg):
    """Sequences cut at max_seq_len (utils/data_utils.py:205-209, :237-244): the oracle on the reference's own truncated inputs."""
    g, batch = load_golden("gen10_truncated")
    out = _run(full_cfg, g, batch)
    assert np.array_equal(out["token_rows"].numpy(), g["token_rows"])
    np.testing.assert_allclose(out["token_logp"].numpy(), g["token_logp"], atol=TOL, rtol=0)
    np.testing.assert_allclose(out["seq_score"].numpy(), g["seq_score"], atol=5 * TOL, rtol=0)
    np.testing.assert_allclose(out["nsp_scores"].numpy(), g["nsp_scores"], atol=TOL, rtol=0)
    assert g["seq_score"][5] == 0 and g["seq_score"][6] == 0          # no masked-copy position left inside 256


def test_config4_nsp_ranking_sample_matches_reference(full_cfg):
    """BASELINE config 4 (val.py:125-131): discriminative masks, NSP probability of 'is the answer' — a 10-candidate sample of the
    100-candidate fixture (the whole round is the GPU test's)."""
    g, batch = load_golden("dis100_default")
    b = {k: v[40:50] for k, v in batch.items()}
    out = _run(full_cfg, g, b)
    np.testing.assert_allclose(out["nsp_scores"].numpy(), g["nsp_scores"][40:50], atol=TOL, rtol=0)
    np.testing.assert_allclose(torch.softmax(out["nsp_scores"], 1)[:, 0].numpy(), g["nsp_prob0"][40:50], atol=TOL, rtol=0)


def test_sweep_fixture_sample_matches_reference(full_cfg):
    """tests/golden/sweep3x100_perturbed.npz (3 rounds x 100 candidates of the bench's own generator, scored by the reference):
    the oracle on 4 candidates of each round, inputs regenerated by unimm_b200.synthetic."""
    from unimm_b200 import synthetic as syn
    from unimm_b200.descriptors import dense_co_mask, dense_text_mask
    g = dict(np.load(os.path.join(os.path.dirname(__file__), "golden", "sweep3x100_perturbed.npz")))
    (feat, loc, mask), rounds = syn.synth_dialog_rounds(int(g["image_id"]), rounds=tuple(int(r) for r in g["round_ids"]))
    assert np.array_equal(np.concatenate([r.tokens for r in rounds]), g["tokens"])
    sd = golden_state_dict(full_cfg, g["weight_seed"], g["perturbed"])
    pick = [0, 33, 66, 99]
    for ri, r in enumerate(rounds):
        desc = torch.from_numpy(r.desc[pick])
        n = len(pick)
        with torch.no_grad():
            out = vo.forward(sd, full_cfg, torch.from_numpy(r.tokens[pick]), torch.from_numpy(feat).expand(n, -1, -1),
                             torch.from_numpy(loc).expand(n, -1, -1), torch.from_numpy(r.segments[pick]), torch.from_numpy(r.positions[pick]),
                             dense_text_mask(desc, 256), torch.from_numpy(mask).expand(n, -1), dense_co_mask(desc, 256).unsqueeze(1).repeat(1, 37, 1),
                             masked_lm_labels=torch.from_numpy(r.labels[pick]))
        np.testing.assert_allclose(out["seq_score"].numpy(), g["seq_score"][ri, pick], atol=5 * TOL, rtol=0)


def test_rank_loss_gradient_oracle_matches_reference_autograd():
    """d neuralNDCG_transposed / d y_pred: the hand-derived chain against ``y_pred.grad`` of the unmodified reference."""
    from oracle import rank_loss as orl
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "rankloss_grad.npz"))
    for p, y, want in zip(z["y_pred"], z["y_true"], z["grad"]):
        got = orl.neural_ndcg_transposed_grad(p, y)
        assert np.abs(got - want).max() < 1e-5 * np.abs(want).max()
    assert np.abs(z["grad"][2][1]).max() == 0.0          # the slate without a relevant option gets no gradient


@pytest.mark.slow
def test_oracle_autograd_matches_reference_backward(full_cfg):
    """The checker of the training step's gradients is torch.autograd over the oracle's forward.  Pin it: on the train6_perturbed batch
    every parameter's gradient (L2 norm, sum, and nine small tensors in full) equals what the UNMODIFIED reference left in ``.grad`` after
    ``(lm + nsp + img).backward()`` (tests/golden/train6_grads.npz, made by tests/golden/make_golden.py)."""
    from test_train_step_cpu import _train_inputs, oracle_losses_and_grads
    from conftest import golden_state_dict
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "train6_grads.npz"))
    g, b, _ = _train_inputs(None)
    sd = golden_state_dict(full_cfg, g["weight_seed"], g["perturbed"])
    losses, grads = oracle_losses_and_grads(full_cfg, sd, b, g, b["tokens"].shape[0], dtype=torch.float32)
    assert abs(sum(losses.values()) - float(z["loss"])) < 1e-4
    gmax = float(z["grad_norm"].max())
    seen = 0
    for name, norm, total, none in zip(z["names"], z["grad_norm"], z["grad_sum"], z["grad_none"]):
        name = str(name)
        if name == "cls.predictions.decoder.weight":
            continue
        mine = grads[name]
        if none:
            assert mine is None, name
            continue
        seen += 1
        assert abs(float(mine.double().norm()) - norm) < 2e-3 * max(norm, 1e-4 * gmax), (name, float(mine.double().norm()), norm)
        assert abs(float(mine.double().sum()) - total) < 2e-3 * max(norm, 1e-4 * gmax) * np.sqrt(mine.numel()), name
        key = "grad__" + name
        if key in z.files:
            want = z[key]
            assert np.abs(mine.numpy() - want).max() < 2e-3 * max(np.abs(want).max(), 1e-6 * gmax), name
    assert seen > 500
