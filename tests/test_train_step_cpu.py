"""CPU: the ORCHESTRATION of the training step (unimm_b200/train_step.py) — layer schedule in reverse, saved activations, fused
Q|K|V views, padded tensors, parameter groups, AdamW ranges — run over a torch-fp64 statement of the device operations
(tests/torch_train_ops.py) and compared with torch.autograd of the oracle.  The CUDA kernels behind the real operations are checked
on the GPU (tests/test_train_step_gpu.py)."""
import json
import os

import numpy as np
import pytest
import torch

from conftest import load_golden
from torch_train_ops import TorchOps


def _train_inputs(n=None):
    g, b = load_golden("train6_perturbed")
    from unimm_b200.descriptors import descriptors_from_masks
    desc = descriptors_from_masks(b["txt_attention_mask"], b["co_attention_mask"])
    sl = slice(0, n)
    batch = {"tokens": b["tokens"][sl], "segments": b["segments"][sl], "positions": b["positions"][sl], "labels": b["mask"][sl],
             "weights": b["weights"][sl], "desc": desc[sl], "next_sentence_label": torch.from_numpy(g["next_sentence_label"])[sl],
             "image_feat": torch.from_numpy(g["image_feat"])[None], "image_loc": torch.from_numpy(g["image_loc"])[None],
             "image_mask": torch.from_numpy(g["image_mask"])[None], "image_label": torch.from_numpy(g["image_label"])[None],
             "image_target": torch.from_numpy(g["image_target"])[None], "seq_image": torch.zeros(b["tokens"][sl].shape[0], dtype=torch.int64),
             "nsp_weight": torch.from_numpy(g["nsp_weight"])}
    return g, b, batch


def oracle_losses_and_grads(cfg, sd, b, g, n, dtype=torch.float64, device="cpu", coeff=(1.0, 1.0, 1.0)):
    """loss and d loss / d parameter from torch.autograd over the oracle's forward (the tied decoder shares the embedding tensor)."""
    from oracle import vilbert_oracle as vo
    p = {k: v.to(device, dtype).clone().requires_grad_() for k, v in sd.items() if k != "cls.predictions.decoder.weight"}
    p["cls.predictions.decoder.weight"] = p["bert.embeddings.word_embeddings.weight"]
    sl = slice(0, n)
    nseq = b["tokens"][sl].shape[0]
    ex = lambda a: torch.from_numpy(a).to(device)[None].expand(nseq, *a.shape)                          # noqa: E731
    o = vo.forward(p, cfg, b["tokens"][sl].to(device), ex(g["image_feat"]), ex(g["image_loc"]), b["segments"][sl].to(device),
                   b["positions"][sl].to(device), b["txt_attention_mask"][sl].to(device), ex(g["image_mask"]),
                   b["co_attention_mask"][sl].to(device), masked_lm_labels=b["mask"][sl].to(device),
                   next_sentence_label=torch.from_numpy(g["next_sentence_label"])[sl].to(device), image_label=ex(g["image_label"]),
                   image_target=ex(g["image_target"]), nsp_weight=torch.from_numpy(g["nsp_weight"]).to(device),
                   lm_weight=b["weights"][sl].to(device), dtype=dtype)
    loss = coeff[0] * o["lm_loss"] + coeff[1] * o["nsp_loss"] + coeff[2] * o["img_loss"]          # train.py:167-168
    names = [k for k in p if k != "cls.predictions.decoder.weight"]
    grads = torch.autograd.grad(loss, [p[k] for k in names], allow_unused=True)
    return {k: float(o[k].detach()) for k in ("lm_loss", "nsp_loss", "img_loss")}, dict(zip(names, grads))


@pytest.mark.slow
def test_train_step_orchestration_matches_autograd_of_the_oracle():
    from oracle import vilbert_oracle as vo   # noqa: F401  (device-independent oracle)
    from unimm_b200.config import tiny_config
    from unimm_b200.train_step import TrainStep
    from unimm_b200.weights import random_state_dict
    cfg = tiny_config()
    sd = random_state_dict(cfg, seed=5, perturbed=True)
    n = 3
    g, b, batch = _train_inputs(n)
    ref_loss, ref_grad = oracle_losses_and_grads(cfg, sd, b, g, n)
    ts = TrainStep(cfg, sd, TorchOps())
    vals = ts.forward_backward(batch)
    for k in ("lm_loss", "nsp_loss", "img_loss"):
        assert abs(vals[k] - ref_loss[k]) < 1e-9, (k, vals[k], ref_loss[k])
    got = ts.grad_dict()
    worst = 0.0
    gmax = max(float(gr.abs().max()) for gr in ref_grad.values() if gr is not None)
    for name, gr in ref_grad.items():
        if gr is None:                                    # parameters the forward never touches
            assert float(got[name].abs().max()) == 0.0, name
            continue
        scale = max(float(gr.abs().max()), 1e-6 * gmax)      # key biases: the exact gradient is 0 (softmax is shift invariant)
        err = float((got[name].double() - gr).abs().max()) / scale
        worst = max(worst, err)
        assert err < 1e-8, (name, err)
    print(f"orchestration vs autograd: worst relative gradient error {worst:.2e} over {len(ref_grad)} tensors")
    # loss coefficients other than 1 (train.py:167-168) and gradient accumulation's 1 / batch_multiply (train.py:451)
    _, ref_c = oracle_losses_and_grads(cfg, sd, b, g, n, coeff=(0.35, 0.65, 0.25))
    tsc = TrainStep(cfg, sd, TorchOps(), lm_coeff=0.7, nsp_coeff=1.3, img_coeff=0.5, batch_multiply=2)
    tsc.forward_backward(batch)
    gotc = tsc.grad_dict()
    for name, gr in ref_c.items():
        if gr is not None:
            assert float((gotc[name].double() - gr).abs().max()) < 1e-8 * max(float(gr.abs().max()), 1e-6 * gmax), name
    # the optimizer: two steps against the restated pytorch_transformers AdamW with the reference's groups
    from oracle import adamw as oa
    lw_path = "/root/reference/config/language_weights.json"
    state, ref_p = {}, {k: v.double().clone() for k, v in sd.items()}
    from unimm_b200.train_step import is_language_weight
    if os.path.exists(lw_path):                           # the group rule against the reference's own list (build container only)
        lw = set(json.load(open(lw_path)))
        for k in sd:
            if k != "cls.predictions.decoder.weight":
                assert (("bert_pretrained." + k) in lw) == is_language_weight(k), k
    ts2 = TrainStep(cfg, sd, TorchOps(), lr=3e-5, image_lr=5e-5, warmup_steps=1, t_total=10)
    for it in range(2):
        ts2.step(batch)
        loss_i, grad_i = oracle_losses_and_grads(cfg, {k: v.float() for k, v in ref_p.items()} if False else ref_p, b, g, n)
        lr_l, lr_v = oa.warmup_linear_nonzero_lr(it, 3e-5, 1, 10), oa.warmup_linear_nonzero_lr(it, 5e-5, 1, 10)
        for k, gr in grad_i.items():
            if gr is None:
                continue
            lr = lr_l if is_language_weight(k) else lr_v
            wd = 0.0 if any(nd in k for nd in oa.NO_DECAY) else 0.01
            oa.adamw_step(ref_p[k], gr, state.setdefault(k, {}), lr, weight_decay=wd)
        ref_p["cls.predictions.decoder.weight"] = ref_p["bert.embeddings.word_embeddings.weight"]
    new = ts2.state_dict()
    for k, v in ref_p.items():
        d = float((new[k].double() - v).abs().max())
        assert d < 1e-6, (k, d)           # state_dict() exports fp32
        moved = float((v - sd[k].double()).abs().max())
        if not any(m in k for m in ("sep_embeddings", "q_dense")):
            assert moved > 0, k


def test_parameter_groups_and_flat_layout():
    from unimm_b200.config import tiny_config
    from unimm_b200.train_step import ParamStore, param_group
    from unimm_b200.weights import param_shapes
    cfg = tiny_config()
    ps = ParamStore(cfg, TorchOps())
    shapes = param_shapes(cfg)
    assert set(ps.entries) == set(shapes) - {"cls.predictions.decoder.weight"}
    # fused Q|K|V spans are gap-free
    a = "bert.encoder.layer.0.attention.self."
    w = ps.span(ps.p, a + "query.weight", a + "value.weight")
    assert tuple(w.shape) == (3 * cfg.hidden_size, cfg.hidden_size)
    b = "bert.encoder.c_layer.0.biattention."
    assert tuple(ps.span(ps.p, b + "query2.weight", b + "value2.weight").shape) == (3 * cfg.bi_hidden_size, cfg.hidden_size)
    assert ps.span(ps.p, b + "query1.bias", b + "value1.bias").numel() == 3 * cfg.bi_hidden_size
    # the substring rule of train.py:323: LayerNorm1/2.weight of the connection layers DO decay, every bias does not
    assert param_group("bert.encoder.c_layer.0.biOutput.LayerNorm1.weight") == 2
    assert param_group("bert.encoder.c_layer.0.biOutput.LayerNorm1.bias") == 3
    assert param_group("bert.encoder.layer.0.output.LayerNorm.weight") == 1
    assert param_group("bert.t_pooler.dense.weight") == 2 and param_group("cls.predictions.bias") == 1
    assert param_group("bert.embeddings.sep_embeddings.weight") == 4
    for name, (off, pad, shp) in ps.entries.items():
        assert off % 64 == 0 and all(p >= s for p, s in zip(pad, shp)), name


@pytest.mark.slow
def test_dense_annotation_step_orchestration():
    """dense_annotation_finetuning.py:253-296: loss = neuralNDCG_transposed(softmax(nsp)[:, 0], relevance) + lm + nsp CE (unweighted), no
    image term — the image head's parameters get no gradient and are left alone by the optimizer."""
    from oracle import rank_loss as orl
    from oracle import vilbert_oracle as vo
    from unimm_b200.config import tiny_config
    from unimm_b200.train_step import TrainStep
    from unimm_b200.weights import random_state_dict
    cfg = tiny_config()
    sd = random_state_dict(cfg, seed=6, perturbed=True)
    n = 4
    g, b, batch = _train_inputs(n)
    rel = torch.tensor([[0.0, 0.4, 1.0, 0.0]])
    batch = dict(batch, gt_relevance=rel, nsp_weight=None)
    p = {k: v.double().clone().requires_grad_() for k, v in sd.items() if k != "cls.predictions.decoder.weight"}
    p["cls.predictions.decoder.weight"] = p["bert.embeddings.word_embeddings.weight"]
    ex = lambda a: torch.from_numpy(a)[None].expand(n, *a.shape)                                          # noqa: E731
    o = vo.forward(p, cfg, b["tokens"][:n], ex(g["image_feat"]), ex(g["image_loc"]), b["segments"][:n], b["positions"][:n],
                   b["txt_attention_mask"][:n], ex(g["image_mask"]), b["co_attention_mask"][:n], masked_lm_labels=b["mask"][:n],
                   next_sentence_label=torch.from_numpy(g["next_sentence_label"])[:n], image_label=ex(g["image_label"]),
                   image_target=ex(g["image_target"]), nsp_weight=None, lm_weight=b["weights"][:n], dtype=torch.float64)
    p0 = torch.softmax(o["nsp_scores"], -1)[:, 0]
    dnd = torch.from_numpy(orl.neural_ndcg_transposed_grad(p0.detach().numpy()[None], rel.numpy()))[0]
    loss = o["lm_loss"] + o["nsp_loss"] + (p0 * dnd).sum()          # the last term's gradient IS the NDCG chain (dnd is a constant)
    names = [k for k in p if k != "cls.predictions.decoder.weight"]
    grads = dict(zip(names, torch.autograd.grad(loss, [p[k] for k in names], allow_unused=True)))
    ts = TrainStep(cfg, sd, TorchOps(), img_coeff=0.0)
    vals = ts.forward_backward(batch)
    want_nd = float(orl.neural_ndcg_transposed(p0.detach().numpy()[None], rel.numpy())[0])
    assert abs(vals["ndcg_loss"] - want_nd) < 1e-6 and abs(vals["lm_loss"] - float(o["lm_loss"].detach())) < 1e-9
    got = ts.grad_dict()
    gmax = max(float(x.abs().max()) for x in grads.values() if x is not None)
    for k, gr in grads.items():
        if gr is None:
            assert float(got[k].abs().max()) == 0.0, k
            continue
        assert float((got[k].double() - gr).abs().max()) < 1e-8 * max(float(gr.abs().max()), 1e-6 * gmax), k
    assert all(grads[k] is None for k in names if k.startswith("cls.imagePredictions."))
    before = {k: v.clone() for k, v in ts.state_dict().items()}
    ts.optimizer_step()
    after = ts.state_dict()
    for k in names:
        changed = not torch.equal(before[k], after[k])
        if grads[k] is None:
            assert not changed, k                                  # no gradient: not even weight decay
        elif float(grads[k].abs().max()) > 1e-6 * gmax:            # (key biases have an analytically zero gradient)
            assert changed, k


@pytest.mark.slow
def test_gradient_accumulation_and_resume():
    """train.py:451-463: loss / batch_multiply, optimizer every batch_multiply-th iteration (and at iteration 0), scheduler every iteration;
    and a checkpoint of the optimizer state resumes bit-identically."""
    from unimm_b200.config import tiny_config
    from unimm_b200.train_step import TrainStep
    from unimm_b200.weights import random_state_dict
    cfg = tiny_config()
    sd = random_state_dict(cfg, seed=5, perturbed=True)
    _, _, batch = _train_inputs(4)
    halves = [{k: (v[:2] if torch.is_tensor(v) and v.shape[:1] == (4,) else v) for k, v in batch.items()},
              {k: (v[2:] if torch.is_tensor(v) and v.shape[:1] == (4,) else v) for k, v in batch.items()}]
    kw = dict(lr=1e-3, image_lr=1e-3, warmup_steps=0)
    acc = TrainStep(cfg, sd, TorchOps(), batch_multiply=2, **kw)
    acc.step(halves[0])                                  # iteration 0 steps on its own (the reference's `or iter_id == 0`)
    assert acc.opt_step == 1 and acc.sched_step == 1
    p_after0 = acc.params.p.clone()
    acc.step(halves[1])                                  # iteration 1: accumulate only
    assert acc.opt_step == 1 and acc.sched_step == 2 and torch.equal(acc.params.p, p_after0)
    acc.step(halves[0])                                  # iteration 2: steps on the sum of iterations 1 and 2
    assert acc.opt_step == 2 and acc.sched_step == 3
    # the same three iterations by hand
    ref = TrainStep(cfg, sd, TorchOps(), batch_multiply=2, **kw)
    ref.forward_backward(halves[0]); ref.optimizer_step()
    ref.forward_backward(halves[1]); g1 = ref.params.g.clone(); ref.sched_step += 1
    ref.forward_backward(halves[0]); ref.params.g.add_(g1); ref.optimizer_step()
    assert float((ref.params.p - acc.params.p).abs().max()) < 1e-12
    # resume: model + optimizer state into a fresh object, one more step on both
    resumed = TrainStep(cfg, acc.state_dict(), TorchOps(), batch_multiply=2, **kw)
    resumed.params.p.copy_(acc.params.p)                 # state_dict() exports fp32; carry the test's fp64 masters over exactly
    resumed.params.refresh_lp()
    resumed.load_optimizer_state_dict(acc.optimizer_state_dict())
    resumed.params.m.copy_(acc.params.m); resumed.params.v.copy_(acc.params.v)
    a, b = acc.step(halves[1]), resumed.step(halves[1])
    a2, b2 = acc.step(halves[0]), resumed.step(halves[0])
    assert a == b and a2 == b2 and resumed.opt_step == acc.opt_step == 3
    assert float((resumed.params.p - acc.params.p).abs().max()) == 0.0
    osd = acc.optimizer_state_dict()
    assert set(osd["exp_avg"]) == set(acc.params.entries) and osd["exp_avg"]["cls.bi_seq_relationship.weight"].shape == (2, 1024)


@pytest.mark.slow
def test_train_step_with_dropout_matches_autograd_under_the_same_masks():
    """Dropout on (p = 0.1 at every nn.Dropout site of the reference): the step draws counter-based keep-masks (site, forward number,
    element index) and regenerates them in the backward; the oracle, given the SAME masks at the reference's dropout call sites,
    must have the same losses and gradients."""
    from torch_train_ops import keep_mask
    from unimm_b200.config import tiny_config
    from unimm_b200.train_step import TrainStep, site_seed
    from unimm_b200.weights import random_state_dict
    from oracle import vilbert_oracle as vo
    cfg = tiny_config()
    sd = random_state_dict(cfg, seed=5, perturbed=True)
    n = 3
    g, b, batch = _train_inputs(n)
    ts = TrainStep(cfg, sd, TorchOps(), dropout=0.1, seed=1234)
    vals = ts.forward_backward(batch)
    sites = []

    def drop(site, x):
        sites.append(site)
        return x * keep_mask((site_seed(1234 + 0, site), 0.1), tuple(x.shape))      # forward number 0

    p = {k: v.double().clone().requires_grad_() for k, v in sd.items() if k != "cls.predictions.decoder.weight"}
    p["cls.predictions.decoder.weight"] = p["bert.embeddings.word_embeddings.weight"]
    ex = lambda a: torch.from_numpy(a)[None].expand(n, *a.shape)                                          # noqa: E731
    o = vo.forward(p, cfg, b["tokens"][:n], ex(g["image_feat"]), ex(g["image_loc"]), b["segments"][:n], b["positions"][:n],
                   b["txt_attention_mask"][:n], ex(g["image_mask"]), b["co_attention_mask"][:n], masked_lm_labels=b["mask"][:n],
                   next_sentence_label=torch.from_numpy(g["next_sentence_label"])[:n], image_label=ex(g["image_label"]),
                   image_target=ex(g["image_target"]), nsp_weight=torch.from_numpy(g["nsp_weight"]), lm_weight=b["weights"][:n],
                   dtype=torch.float64, drop=drop)
    assert len(sites) == len(set(sites)) == 2 + 3 * (cfg.num_hidden_layers + cfg.v_num_hidden_layers) + 6 * len(cfg.v_biattention_id) + 1
    for k in ("lm_loss", "nsp_loss", "img_loss"):
        assert abs(vals[k] - float(o[k].detach())) < 1e-9, k
    names = [k for k in p if k != "cls.predictions.decoder.weight"]
    grads = dict(zip(names, torch.autograd.grad(o["lm_loss"] + o["nsp_loss"] + o["img_loss"], [p[k] for k in names], allow_unused=True)))
    got = ts.grad_dict()
    gmax = max(float(x.abs().max()) for x in grads.values() if x is not None)
    for k, gr in grads.items():
        if gr is not None:
            assert float((got[k].double() - gr).abs().max()) < 1e-8 * max(float(gr.abs().max()), 1e-6 * gmax), k
    # the next forward draws different masks; dropout off reproduces the eval-mode losses
    assert ts.forward_backward(batch)["lm_loss"] != vals["lm_loss"]
