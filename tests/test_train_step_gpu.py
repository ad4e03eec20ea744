"""GPU: the training step — the attention backward and the small training kernels against torch.autograd in
fp64 on the same 16-bit operands, then the whole step (forward + three losses + backward + AdamW, unimm_b200/train_step.py over the
CUDA kernels) against autograd of the oracle and the reference-made loss values of tests/golden/train6_perturbed.npz."""
import math

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

if not torch.cuda.is_available():
    pytest.skip("needs a CUDA device", allow_module_level=True)

from conftest import golden_state_dict, load_golden  # noqa: E402
from test_train_step_cpu import _train_inputs, oracle_losses_and_grads  # noqa: E402
from torch_train_ops import TorchOps  # noqa: E402

from unimm_b200.descriptors import dense_co_mask, dense_text_mask  # noqa: E402
from unimm_b200.train_ops import MASK_CO_INTERVAL, MASK_KEY_VECTOR, MASK_TEXT_SELF, DeviceOps  # noqa: E402

DEV = "cuda:0"


@pytest.mark.parametrize("precision,ulp", [("fp16", 2.0 ** -11), ("bf16", 2.0 ** -8)])
def test_ffn1_pre_activation_kept_as_16_bit_values(precision, ulp, monkeypatch):
    """The FFN-1 forward keeps GELU's input for the backward.  Default: as 16-bit values (the epilogue's pre_act_lp variant; the
    backward's pass over dY reads them, cast_colsum_kernel<2>); UNIMM_PRE16=0: as fp32.  Same activation either way, the stored value
    within one rounding, and the gradients of the projection in front of the GELU agree to the operand rounding."""
    from unimm_b200.train_ops import ACT_GELU
    g = torch.Generator().manual_seed(11)
    M, K, N = 4500, 128, 256
    dt = torch.float16 if precision == "fp16" else torch.bfloat16
    x16 = torch.randn(M, K, generator=g).to(dt).to(DEV)
    w16 = (torch.randn(N, K, generator=g) * 0.2).to(dt).to(DEV)
    bias = torch.randn(N, generator=g).to(DEV)
    dy = (torch.randn(M, N, generator=g) * 1e-3).to(DEV)
    ops16 = DeviceOps(DEV, precision)
    monkeypatch.setenv("UNIMM_PRE16", "0")
    ops32 = DeviceOps(DEV, precision)
    assert ops16.pre16 and not ops32.pre16
    t16, a16 = ops16.linear(x16, w16, bias, act=ACT_GELU, want16=True, pre_act32=True)
    t32, a32 = ops32.linear(x16, w16, bias, act=ACT_GELU, want16=True, pre_act32=True)
    assert t16.dtype == dt and t32.dtype == torch.float32
    want = x16.double() @ w16.double().t() + bias.double()
    assert (t32.double() - want).abs().max().item() < 1e-4 * want.abs().max().item()
    assert ((t16.double() - t32.double()).abs() <= ulp * t32.double().abs() + 1e-7).all()
    assert torch.equal(a16, a32)
    out = []
    for ops, t in ((ops16, t16), (ops32, t32)):
        gw, gb = torch.empty(N, K, device=DEV), torch.empty(N, device=DEV)
        dx = ops.linear_backward(dy.clone(), x16, w16, gw, gb, gelu_t=t)
        out.append((dx, gw, gb))
    torch.cuda.synchronize()
    td = t32.double()
    dpre = dy.double() * (0.5 * (1 + torch.erf(td / math.sqrt(2))) + td * torch.exp(-0.5 * td * td) / math.sqrt(2 * math.pi))
    ref = (dpre @ w16.double(), dpre.t() @ x16.double(), dpre.sum(0))
    for name, a, b, r in zip(("dX", "dW", "db"), out[0], out[1], ref):
        e16, e32 = ((v.double() - r).abs().max().item() / r.abs().max().item() for v in (a, b))
        print(f"[{precision}] {name}: 16-bit pre-activation {e16:.2e}, fp32 pre-activation {e32:.2e} (of the largest reference value)")
        assert e16 < (3e-3 if precision == "fp16" else 2e-2) and e32 < (3e-3 if precision == "fp16" else 2e-2)


def _descs():
    # generative: (mode 0, ctx, L, last_len); discriminative: (1, 0, L, 0); one truncated sequence (L + last_len > S)
    return torch.tensor([[0, 40, 47, 7], [1, 0, 93, 0], [0, 150, 153, 3], [0, 247, 252, 5], [1, 0, 256, 0]], dtype=torch.int32)


@pytest.mark.parametrize("precision,tol", [("fp16", 6e-3), ("bf16", 3e-2)])
@pytest.mark.parametrize("case", ["text_self", "image_self", "text_over_image", "image_over_text"])
def test_attention_backward_matches_autograd(case, precision, tol):
    ops = DeviceOps(DEV, precision)
    ref = TorchOps()
    g = torch.Generator().manual_seed(11)
    desc = _descs()
    B = desc.shape[0]
    S, R = 256, 37
    key_mask = torch.ones(B, R)
    key_mask[1, 30:] = 0
    key_mask[3, 5:] = 0
    heads, D, Sq, Skv, kind = {"text_self": (12, 64, S, S, MASK_TEXT_SELF), "image_self": (8, 128, R, R, MASK_KEY_VECTOR),
                               "text_over_image": (8, 128, S, R, MASK_KEY_VECTOR), "image_over_text": (8, 128, R, S, MASK_CO_INTERVAL)}[case]
    H = heads * D
    rnd = lambda *s: torch.randn(*s, generator=g).to(ops.lp_dtype)                                       # noqa: E731
    q, k, v = rnd(B * Sq, H), rnd(B * Skv, H), rnd(B * Skv, H)
    dO = torch.randn(B * Sq, H, generator=g) * 3e-5                     # gradient-sized values: far below fp16's normal range
    d_desc = desc.to(DEV) if kind != MASK_KEY_VECTOR else None
    d_km = key_mask.to(DEV) if kind == MASK_KEY_VECTOR else None
    o, lse = ops.attention(q.to(DEV), k.to(DEV), v.to(DEV), B, heads, D, Sq, Skv, kind, d_desc, d_km)
    o_ref, lse_ref = ref.attention(q.double(), k.double(), v.double(), B, heads, D, Sq, Skv, kind, desc, key_mask)
    # rows that may attend nothing (padding) are unconstrained (the reference leaves a plain softmax there; nothing reads them)
    if kind == MASK_TEXT_SELF:
        valid = dense_text_mask(desc, S).any(-1).reshape(-1)
    else:
        valid = torch.ones(B * Sq, dtype=torch.bool)
    err_o = (o.float().cpu().double() - o_ref)[valid].abs().max().item()
    lse_v = valid.view(B, 1, Sq).expand(B, heads, Sq)
    err_l = (lse.cpu().double() - lse_ref)[lse_v].abs().max().item()
    print(f"[{precision}] {case}: forward |err| {err_o:.2e}, lse |err| {err_l:.2e}")
    assert err_o < (4e-3 if precision == "fp16" else 3e-2) and err_l < 2e-3
    dO = dO * valid[:, None]                                            # no gradient flows into padding rows
    dq, dk, dv = ops.empty32(B * Sq, H), ops.empty32(B * Skv, H), ops.empty32(B * Skv, H)
    ops.attention_backward(q.to(DEV), k.to(DEV), v.to(DEV), o, lse, dO.to(DEV), B, heads, D, Sq, Skv, kind, d_desc, d_km, dq, dk, dv)
    rq, rk, rv = torch.empty(B * Sq, H, dtype=torch.float64), torch.empty(B * Skv, H, dtype=torch.float64), torch.empty(B * Skv, H, dtype=torch.float64)
    ref.attention_backward(q.double(), k.double(), v.double(), None, None, dO.double(), B, heads, D, Sq, Skv, kind, desc, key_mask, rq, rk, rv)
    for name, mine, want in (("dq", dq, rq), ("dk", dk, rk), ("dv", dv, rv)):
        e = (mine.cpu().double() - want).abs().max().item() / want.abs().max().item()
        print(f"[{precision}] {case} {name}: max |err| / max |ref| = {e:.3e}")
        assert e < tol, name
        assert torch.isfinite(mine).all()


def test_small_training_kernels():
    ops, ref = DeviceOps(DEV, "fp16"), TorchOps()
    g = torch.Generator().manual_seed(3)
    # embedding sum and its scatter-add backward
    V, H, rows = 500, 768, 300
    word, pos_e, ty, ext = (torch.randn(n, H, generator=g) for n in (V, 64, 2, 10))
    ids, pos, seg = torch.randint(0, V, (rows,), generator=g), torch.randint(0, 64, (rows,), generator=g), torch.randint(0, 12, (rows,), generator=g)
    out = ops.embed_text_sum(ids.to(DEV), seg.to(DEV), pos.to(DEV), word.to(DEV), pos_e.to(DEV), ty.to(DEV), ext.to(DEV), 2)
    want = ref.embed_text_sum(ids, seg, pos, word.double(), pos_e.double(), ty.double(), ext.double(), 2)
    assert (out.cpu().double() - want).abs().max().item() < 1e-5
    d = torch.randn(rows, H, generator=g)
    d[::3] = 0
    gw, gp, gt, ge = (torch.zeros(n, H, device=DEV) for n in (V, 64, 2, 10))
    ops.embed_text_backward(d.to(DEV), ids.to(DEV), seg.to(DEV), pos.to(DEV), gw, gp, gt, ge, 2)
    rw, rp, rt, re_ = (torch.zeros(n, H, dtype=torch.float64) for n in (V, 64, 2, 10))
    ref.embed_text_backward(d.double(), ids, seg, pos, rw, rp, rt, re_, 2)
    for a, b in ((gw, rw), (gp, rp), (gt, rt), (ge, re_)):
        assert (a.cpu().double() - b).abs().max().item() < 2e-4 * max(1.0, b.abs().max().item())
    # GELU forward (both outputs), element-wise helpers, gather / scatter
    t = torch.randn(64, 1024, generator=g) * 2
    g32, g16 = ops.gelu(t.to(DEV), want32=True, want16=True)
    wantg = ref.gelu(t.double(), want32=True)[0]
    assert (g32.cpu().double() - wantg).abs().max().item() < 1e-6 and (g16.float().cpu().double() - wantg).abs().max().item() < 4e-3
    a, b = torch.randn(1000, generator=g), torch.randn(1000, generator=g)
    assert torch.equal(ops.mul(a.to(DEV), b.to(DEV)).cpu(), a * b)
    assert torch.equal(ops.relu_backward(a.to(DEV).clone(), b.to(DEV)).cpu(), a * (b > 0))
    src = torch.randn(50, 768, generator=g)
    idx = torch.randperm(50, generator=g)[:20].to(torch.int32)
    assert torch.equal(ops.gather_rows(src.to(DEV), idx.to(DEV)).cpu(), src[idx.long()])
    dst = torch.zeros(50, 768, device=DEV)
    ops.scatter_add_rows(src[:20].to(DEV), idx.to(DEV), dst)
    assert torch.equal(dst.cpu()[idx.long()], src[:20])
    # column sums inside linear_backward at a row count that spans many slabs
    dy = torch.randn(5000, 128, generator=g) * 1e-4
    x = torch.randn(5000, 64, generator=g).to(torch.float16)
    w = torch.randn(128, 64, generator=g).to(torch.float16)
    gwt, gb = torch.empty(128, 64, device=DEV), torch.empty(128, device=DEV)
    dx = ops.linear_backward(dy.to(DEV), x.to(DEV), w.to(DEV), gwt, gb)
    assert (gb.cpu().double() - dy.double().sum(0)).abs().max().item() < 1e-6
    assert (dx.cpu().double() - dy.double() @ w.double()).abs().max().item() < 3e-3 * (dy.double() @ w.double()).abs().max().item()
    acc = torch.ones(5000, 64, device=DEV)
    ops.linear_backward(dy.to(DEV), x.to(DEV), w.to(DEV), gwt, gb, dx_accum=acc)
    assert (acc.cpu() - 1.0 - dx.cpu()).abs().max().item() < 1e-6
    # NSP cross entropy, image KL
    logits = torch.randn(37, 2, generator=g)
    y = torch.randint(0, 2, (37,), generator=g)
    nw = torch.tensor([2.0, 5.0])
    loss, dl = ops.nsp_ce(logits.to(DEV), y.to(DEV), nw.to(DEV), 0.7)
    rl, rd = ref.nsp_ce(logits.double(), y, nw.double(), 0.7)
    assert abs(loss.item() - rl.item()) < 1e-5 and (dl.cpu().double() - rd).abs().max().item() < 1e-6
    C, ld, n_img = 1601, 1664, 3
    vl = torch.randn(3 * 37, ld, generator=g)
    tgt = torch.rand(n_img * 37, C, generator=g) ** 8
    tgt /= tgt.sum(-1, keepdim=True)
    trow = torch.randint(0, n_img * 37, (3 * 37,), generator=g).to(torch.int32)
    il = torch.where(torch.rand(3 * 37, generator=g) < 0.2, 1, -1)
    loss, dv = ops.image_kl(vl.to(DEV), C, tgt.to(DEV), trow.to(DEV), il.to(DEV), 1.3)
    rl, rd = ref.image_kl(vl.double(), C, tgt.double(), trow, il, 1.3)
    assert abs(loss.item() - rl.item()) < 1e-4 * abs(rl.item()) and (dv.cpu().double() - rd).abs().max().item() < 1e-6
    # AdamW against the restated pytorch_transformers step
    from oracle import adamw as oa
    p, gr = torch.randn(10000, generator=g), torch.randn(10000, generator=g) * 1e-3
    state, rp = {}, p.double().clone()
    dp, dm, dv_, d16 = p.to(DEV).clone(), torch.zeros(10000, device=DEV), torch.zeros(10000, device=DEV), torch.empty(10000, device=DEV, dtype=torch.float16)
    for it in range(1, 4):
        oa.adamw_step(rp, gr.double(), state, 2e-5, weight_decay=0.01)
        ops.adamw(dp, gr.to(DEV), dm, dv_, 2e-5, 0.9, 0.999, 1e-6, 0.01, it, True, 1.0, d16)
    assert (dp.cpu().double() - rp).abs().max().item() < 1e-6 and torch.equal(d16.cpu(), dp.cpu().half())


def _compare_grads(got, ref_grad, tol_max, tol_l2, label):
    """Per tensor: max |err| relative to the tensor's largest gradient, and the relative L2 error.  Tensors whose gradient is
    (analytically) zero or negligible next to the largest one in the model are measured against 1e-3 of that."""
    gmax = max(float(gr.abs().max()) for gr in ref_grad.values() if gr is not None)
    rows = []
    for name, gr in ref_grad.items():
        if gr is None:
            assert float(got[name].abs().max()) == 0.0, name
            continue
        assert torch.isfinite(got[name]).all(), name
        d = got[name].double() - gr.double().cpu()
        scale = max(float(gr.abs().max()), 1e-3 * gmax)
        l2 = float(d.norm()) / max(float(gr.double().norm()), 1e-3 * gmax * math.sqrt(gr.numel()))
        rows.append((float(d.abs().max()) / scale, l2, name))
    rows.sort(reverse=True)
    for e, l2, name in rows[:6]:
        print(f"[{label}]   max-err {e:.3e}  l2-err {l2:.3e}  {name}")
    worst_l2 = max(r[1] for r in rows)
    print(f"[{label}] {len(rows)} gradient tensors: worst max-err {rows[0][0]:.3e} ({rows[0][2]}), worst relative L2 error {worst_l2:.3e}, "
          f"median max-err {sorted(r[0] for r in rows)[len(rows) // 2]:.3e}")
    assert rows[0][0] < tol_max, rows[0]
    assert worst_l2 < tol_l2, max(rows, key=lambda r: r[1])


@pytest.mark.parametrize("precision,tol", [("fp16", (2e-2, 1e-2)), ("bf16", (1e-1, 6e-2))])
def test_train_step_tiny_matches_autograd_of_the_oracle(precision, tol):
    from unimm_b200.config import tiny_config
    from unimm_b200.train_step import TrainStep
    from unimm_b200.weights import random_state_dict
    cfg = tiny_config()
    sd = random_state_dict(cfg, seed=5, perturbed=True)
    n = 6
    g, b, batch = _train_inputs(n)
    ref_loss, ref_grad = oracle_losses_and_grads(cfg, sd, b, g, n)
    ts = TrainStep(cfg, sd, DeviceOps(DEV, precision))
    vals = ts.forward_backward(batch)
    print(f"[{precision}] tiny losses {vals} vs {ref_loss}")
    for k in ("lm_loss", "nsp_loss", "img_loss"):
        assert abs(vals[k] - ref_loss[k]) < (5e-3 if precision == "fp16" else 3e-2), (k, vals[k], ref_loss[k])
    _compare_grads(ts.grad_dict(), ref_grad, tol[0], tol[1], precision + " tiny")
    # two optimizer steps move the loss down on the same batch and keep every parameter finite
    first = vals["loss"]
    ts2 = TrainStep(cfg, sd, DeviceOps(DEV, precision), lr=1e-3, image_lr=1e-3, warmup_steps=0)
    losses = [ts2.step(batch)["loss"] for _ in range(4)]
    print(f"[{precision}] loss over 4 steps at lr 1e-3: {losses}")
    assert abs(losses[0] - first) < 1e-2 and losses[-1] < losses[0]
    assert all(torch.isfinite(v).all() for v in ts2.state_dict().values())


def test_train_step_full_config_matches_reference_losses_and_autograd(full_cfg):
    """12 + 6 + 6 layers, the reference-made batch of train6_perturbed: the three losses against the reference's own values, every
    gradient against autograd of the oracle (fp32 on the host cores)."""
    from unimm_b200.train_step import TrainStep
    g, b, batch = _train_inputs(None)
    sd = golden_state_dict(full_cfg, g["weight_seed"], g["perturbed"])
    ts = TrainStep(full_cfg, sd, DeviceOps(DEV, "fp16"))
    vals = ts.forward_backward(batch)
    print(f"[fp16] full-config losses {vals} vs reference lm {g['lm_loss'].item():.6f} img {g['img_loss'].item():.6f} nsp {g['nsp_loss'].item():.6f}")
    assert abs(vals["lm_loss"] - g["lm_loss"].item()) < 2e-2
    assert abs(vals["img_loss"] - g["img_loss"].item()) < 2e-2
    assert abs(vals["nsp_loss"] - g["nsp_loss"].item()) < 2e-2
    n = b["tokens"].shape[0]
    _, ref_grad = oracle_losses_and_grads(full_cfg, sd, b, g, n, dtype=torch.float32)
    got = ts.grad_dict()
    _compare_grads(got, ref_grad, 1e-1, 3e-2, "fp16 full config")
    # and directly against what the UNMODIFIED reference left in .grad after loss.backward() on this batch (tests/golden/train6_grads.npz)
    import os
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "train6_grads.npz"))
    gmax, worst = float(z["grad_norm"].max()), 0.0
    for name, norm, none in zip(z["names"], z["grad_norm"], z["grad_none"]):
        name = str(name)
        if name == "cls.predictions.decoder.weight" or none:
            continue
        worst = max(worst, abs(float(got[name].double().norm()) - norm) / max(norm, 1e-3 * gmax))
        key = "grad__" + name
        if key in z.files:
            want = z[key]
            e = np.abs(got[name].numpy() - want).max() / max(np.abs(want).max(), 1e-9)
            print(f"[fp16 full config] vs the reference's own .grad: {name}: max |err| / max |grad| = {e:.3e}")
            assert e < 1e-1, name
    print(f"[fp16 full config] gradient L2 norms vs the reference's: worst relative difference {worst:.3e}")
    assert worst < 5e-2


@pytest.mark.parametrize("name", ["ft100gen_perturbed", "ft100dis_perturbed"])
def test_dense_annotation_training_step_at_size(full_cfg, name):
    """BASELINE config 5 as a TRAINING step (dense_annotation_finetuning.py:253-296): the 100 options of one annotated round, loss =
    neuralNDCG_transposed + lm + NSP CE (no image term) — the loss values against the reference's, then backward + AdamW; the image
    head, which gets no gradient, is not touched by the optimizer."""
    from unimm_b200.descriptors import descriptors_from_masks
    from unimm_b200.train_step import TrainStep
    g, b = load_golden(name)
    n = b["tokens"].shape[0]
    sd = golden_state_dict(full_cfg, g["weight_seed"], g["perturbed"])
    batch = {"tokens": b["tokens"], "segments": b["segments"], "positions": b["positions"], "labels": b["mask"], "weights": b["weights"],
             "desc": descriptors_from_masks(b["txt_attention_mask"], b["co_attention_mask"]),
             "next_sentence_label": torch.from_numpy(g["next_sentence_label"]), "image_feat": torch.from_numpy(g["image_feat"])[None],
             "image_loc": torch.from_numpy(g["image_loc"])[None], "image_mask": torch.from_numpy(g["image_mask"])[None],
             "image_label": torch.from_numpy(g["image_label"])[None], "image_target": torch.from_numpy(g["image_target"])[None],
             "seq_image": torch.zeros(n, dtype=torch.int64), "gt_relevance": torch.from_numpy(g["relevance"]).view(1, n)}
    ts = TrainStep(full_cfg, sd, DeviceOps(DEV, "fp16"), img_coeff=0.0, lr=5e-5, image_lr=5e-5, warmup_steps=0)
    vals = ts.step(batch)
    print(f"[fp16] {name}: lm {vals['lm_loss']:.6f}/{g['lm_loss'].item():.6f} nsp {vals['nsp_loss']:.6f}/{float(g['nsp_ce_unweighted']):.6f} "
          f"neuralNDCG {vals['ndcg_loss']:.6f}/{float(g['neural_ndcg_loss']):.6f} total {vals['loss']:.6f}/{float(g['total_loss']):.6f}")
    assert abs(vals["lm_loss"] - g["lm_loss"].item()) < 2e-2 and abs(vals["nsp_loss"] - float(g["nsp_ce_unweighted"])) < 2e-2
    assert abs(vals["ndcg_loss"] - float(g["neural_ndcg_loss"])) < 2e-2 and abs(vals["loss"] - float(g["total_loss"])) < 6e-2
    new = ts.state_dict()
    assert all(torch.isfinite(v).all() for v in new.values())
    for k, v in new.items():
        if k.startswith("cls.imagePredictions.") or "q_dense" in k or "sep_embeddings" in k:
            assert torch.equal(v, sd[k].float()), k
    assert not torch.equal(new["cls.bi_seq_relationship.weight"], sd["cls.bi_seq_relationship.weight"].float())
    second = ts.step(batch)
    print(f"[fp16] {name}: total loss after one step at lr 5e-5: {second['loss']:.6f}")
    assert second["loss"] < vals["loss"]


@pytest.mark.parametrize("precision", ["fp16", "bf16"])
def test_config3_training_step_at_size(full_cfg, precision):
    """BASELINE config 3 at its stated size as a whole TRAINING step (train.py:53-92, :445-463): 240 sequences = 40 images x (1 positive +
    5 negatives) from the reference's own encoders, one feature / target block per image (``seq_image``): the three losses against the
    reference's values, then backward + AdamW — every updated parameter finite, the loss on the same batch lower after the step."""
    from conftest import load_golden_multi_image
    from unimm_b200.descriptors import descriptors_from_masks
    from unimm_b200.train_step import TrainStep
    g, b, im = load_golden_multi_image("train240_perturbed")
    sd = golden_state_dict(full_cfg, g["weight_seed"], g["perturbed"])
    batch = {"tokens": b["tokens"], "segments": b["segments"], "positions": b["positions"], "labels": b["mask"], "weights": b["weights"],
             "desc": descriptors_from_masks(b["txt_attention_mask"], b["co_attention_mask"]), "seq_image": b["seq_image"],
             "next_sentence_label": torch.from_numpy(g["next_sentence_label"]), "nsp_weight": torch.from_numpy(g["nsp_weight"]), **im}
    ts = TrainStep(full_cfg, sd, DeviceOps(DEV, precision), lr=1e-4, image_lr=1e-4, warmup_steps=0)
    vals = ts.step(batch)
    print(f"[{precision}] config 3 training step, B = 240: lm {vals['lm_loss']:.6f}/{g['lm_loss'].item():.6f} img {vals['img_loss']:.6f}/"
          f"{g['img_loss'].item():.6f} nsp {vals['nsp_loss']:.6f}/{g['nsp_loss'].item():.6f}")
    tol = 2e-2 if precision == "fp16" else 3e-2
    assert abs(vals["lm_loss"] - g["lm_loss"].item()) < tol and abs(vals["img_loss"] - g["img_loss"].item()) < tol
    assert abs(vals["nsp_loss"] - g["nsp_loss"].item()) < tol
    gn = ts.params.g[:ts.params.group_range[3][1]]
    assert torch.isfinite(gn).all() and float(gn.abs().max()) > 0
    again = ts.step(batch)
    print(f"[{precision}] loss {vals['loss']:.5f} -> {again['loss']:.5f} after one AdamW step at lr 1e-4")
    assert again["loss"] < vals["loss"]
    assert all(torch.isfinite(v).all() for v in ts.state_dict().values())


def test_dropout_kernels_regenerate_the_reference_masks():
    """Counter-based dropout: the element-wise kernel, the GEMM epilogue, the cast pass of the linear backward and the three attention
    kernels all draw mask(seed, i) = tests/torch_train_ops.py::keep_mask — checked against torch with that mask applied."""
    from torch_train_ops import keep_mask
    ops, ref = DeviceOps(DEV, "fp16"), TorchOps()
    g = torch.Generator().manual_seed(21)
    drop = (0x9E3779B1, 0.1)
    x = torch.randn(300, 768, generator=g)
    y32, y16 = ops.dropout(x.to(DEV), drop)
    want = x.double() * keep_mask(drop, x.shape)
    assert (y32.cpu().double() - want).abs().max().item() < 1e-6 and abs(float((y32 == 0).float().mean()) - 0.1) < 0.01
    assert (y16.float().cpu().double() - want).abs().max().item() < 4e-3
    dy = torch.randn(300, 768, generator=g)
    assert (ops.dropout_backward(dy.to(DEV).clone(), drop).cpu().double() - dy.double() * keep_mask(drop, dy.shape)).abs().max().item() < 1e-6
    # projection with output dropout + residual, and its backward
    M, N, K = 1000, 768, 1024
    a, w = torch.randn(M, K, generator=g).half(), (0.03 * torch.randn(N, K, generator=g)).half()
    bias, res = torch.randn(N, generator=g), torch.randn(M, N, generator=g)
    out, _ = ops.linear(a.to(DEV), w.to(DEV), bias.to(DEV), residual=res.to(DEV), drop=drop)
    want = (a.double() @ w.double().t() + bias.double()) * keep_mask(drop, (M, N)) + res.double()
    assert (out.cpu().double() - want).abs().max().item() < 2e-2
    dyl = torch.randn(M, N, generator=g) * 1e-4
    gw, gb = torch.empty(N, K, device=DEV), torch.empty(N, device=DEV)
    dx = ops.linear_backward(dyl.to(DEV), a.to(DEV), w.to(DEV), gw, gb, drop=drop)
    dym = dyl.double() * keep_mask(drop, (M, N))
    for name, mine, wantg in (("dx", dx, dym @ w.double()), ("dw", gw, dym.t() @ a.double()), ("db", gb, dym.sum(0))):
        e = (mine.cpu().double() - wantg).abs().max().item() / wantg.abs().max().item()
        assert e < 4e-3, (name, e)
    # attention with dropout on the probabilities: forward and backward, two mask families
    desc = _descs()
    B, S, R = desc.shape[0], 256, 37
    km = torch.ones(B, R)
    km[1, 30:] = 0
    for heads, D, Sq, Skv, kind in ((12, 64, S, S, MASK_TEXT_SELF), (8, 128, R, S, MASK_CO_INTERVAL), (8, 128, S, R, MASK_KEY_VECTOR)):
        H = heads * D
        q, k, v = (torch.randn(n, H, generator=g).half() for n in (B * Sq, B * Skv, B * Skv))
        d_desc = desc.to(DEV) if kind != MASK_KEY_VECTOR else None
        d_km = km.to(DEV) if kind == MASK_KEY_VECTOR else None
        o, lse = ops.attention(q.to(DEV), k.to(DEV), v.to(DEV), B, heads, D, Sq, Skv, kind, d_desc, d_km, drop=drop)
        o_ref, _ = ref.attention(q.double(), k.double(), v.double(), B, heads, D, Sq, Skv, kind, desc, km, drop=drop)
        valid = dense_text_mask(desc, S).any(-1).reshape(-1) if kind == MASK_TEXT_SELF else torch.ones(B * Sq, dtype=torch.bool)
        assert (o.float().cpu().double() - o_ref)[valid].abs().max().item() < 6e-3
        dO = torch.randn(B * Sq, H, generator=g) * 3e-5 * valid[:, None]
        dq, dk, dv = ops.empty32(B * Sq, H), ops.empty32(B * Skv, H), ops.empty32(B * Skv, H)
        ops.attention_backward(q.to(DEV), k.to(DEV), v.to(DEV), o, lse, dO.to(DEV), B, heads, D, Sq, Skv, kind, d_desc, d_km, dq, dk, dv, drop=drop)
        rq, rk, rv = (torch.empty(n, H, dtype=torch.float64) for n in (B * Sq, B * Skv, B * Skv))
        ref.attention_backward(q.double(), k.double(), v.double(), None, None, dO.double(), B, heads, D, Sq, Skv, kind, desc, km, rq, rk, rv, drop=drop)
        for name, mine, wantg in (("dq", dq, rq), ("dk", dk, rk), ("dv", dv, rv)):
            e = (mine.cpu().double() - wantg).abs().max().item() / wantg.abs().max().item()
            print(f"attention with dropout, mask kind {kind} {name}: max |err| / max |ref| = {e:.3e}")
            assert e < 6e-3, (kind, name)


def test_train_step_with_dropout_matches_autograd_under_the_same_masks_gpu():
    """The whole step with p = 0.1 at every nn.Dropout site of the reference, against autograd of the oracle given the same masks."""
    from torch_train_ops import keep_mask
    from oracle import vilbert_oracle as vo
    from unimm_b200.config import tiny_config
    from unimm_b200.train_step import TrainStep, site_seed
    from unimm_b200.weights import random_state_dict
    cfg = tiny_config()
    sd = random_state_dict(cfg, seed=5, perturbed=True)
    n = 4
    g, b, batch = _train_inputs(n)
    ts = TrainStep(cfg, sd, DeviceOps(DEV, "fp16"), dropout=0.1, seed=77)
    vals = ts.forward_backward(batch)
    p = {k: v.double().clone().requires_grad_() for k, v in sd.items() if k != "cls.predictions.decoder.weight"}
    p["cls.predictions.decoder.weight"] = p["bert.embeddings.word_embeddings.weight"]
    ex = lambda a: torch.from_numpy(a)[None].expand(n, *a.shape)                                          # noqa: E731
    o = vo.forward(p, cfg, b["tokens"][:n], ex(g["image_feat"]), ex(g["image_loc"]), b["segments"][:n], b["positions"][:n],
                   b["txt_attention_mask"][:n], ex(g["image_mask"]), b["co_attention_mask"][:n], masked_lm_labels=b["mask"][:n],
                   next_sentence_label=torch.from_numpy(g["next_sentence_label"])[:n], image_label=ex(g["image_label"]),
                   image_target=ex(g["image_target"]), nsp_weight=torch.from_numpy(g["nsp_weight"]), lm_weight=b["weights"][:n],
                   dtype=torch.float64, drop=lambda site, x: x * keep_mask((site_seed(77, site), 0.1), tuple(x.shape)))
    print(f"[fp16, dropout 0.1] losses {vals} vs oracle with the same masks lm {float(o['lm_loss']):.6f} nsp {float(o['nsp_loss']):.6f} "
          f"img {float(o['img_loss']):.6f}")
    for k in ("lm_loss", "nsp_loss", "img_loss"):
        assert abs(vals[k] - float(o[k].detach())) < 5e-3, k
    names = [k for k in p if k != "cls.predictions.decoder.weight"]
    grads = dict(zip(names, torch.autograd.grad(o["lm_loss"] + o["nsp_loss"] + o["img_loss"], [p[k] for k in names], allow_unused=True)))
    _compare_grads(ts.grad_dict(), grads, 3e-2, 1e-2, "fp16 tiny, dropout 0.1")
    ts_off = TrainStep(cfg, sd, DeviceOps(DEV, "fp16"))
    assert abs(ts_off.forward_backward(batch)["lm_loss"] - vals["lm_loss"]) > 1e-4            # the masks really were applied
