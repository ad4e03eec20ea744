"""GPU: the drop-in driven by the reference's OWN callers: ``train.forward`` (train.py:30-177) in its
evaluation and training branches, and the chunk loop + metrics of ``val_lm.visdial_evaluate`` (val_lm.py:40-191), executed
unmodified from a reference checkout over ``unimm_b200.VisualDialogEncoder`` — against the reference-made goldens and against
the unmodified reference model running as eager PyTorch on the same GPU."""
import numpy as np
import pytest
import torch

from conftest import golden_state_dict, load_golden
from ref_callers import build_reference_encoder, import_reference, reference_root

pytestmark = pytest.mark.gpu

if not torch.cuda.is_available():
    pytest.skip("needs a CUDA device", allow_module_level=True)
REF_ROOT = reference_root()
if REF_ROOT is None:
    pytest.skip("no reference checkout (run scripts/make_ref_copy.py in the build container)", allow_module_level=True)

from unimm_b200.visual_dialog_encoder import VisualDialogEncoder  # noqa: E402

REF = import_reference(REF_ROOT)


def ours(cfg, g, precision="fp32", max_sequences=32):
    enc = VisualDialogEncoder(cfg, precision=precision, max_sequences=max_sequences)
    sd = golden_state_dict(cfg, g["weight_seed"], g["perturbed"])
    enc.load_state_dict({"bert_pretrained." + k: v for k, v in sd.items()}, strict=True)
    return enc


def loader_batch(b, extra=None):
    """The reference loader's layout: a leading image dimension over [options, ...] tensors (dataloader_visdial.py:437-457)."""
    n = b["tokens"].shape[0]
    out = {k: b[k].unsqueeze(0) for k in ("tokens", "segments", "positions", "weights", "sep_indices", "mask", "txt_attention_mask",
                                          "co_attention_mask", "image_feat", "image_loc", "image_mask")}
    out["hist_len"] = torch.zeros(1, n, dtype=torch.long)
    out.update(extra or {})
    return out


def test_train_forward_evaluation_branch(full_cfg):
    """val_lm.py:121: forward(encoder, item, params, output_nsp_scores=True, output_lm_scores=True, evaluation=True)."""
    import torch.nn.functional as F
    g, b = load_golden("gen8_perturbed")
    enc = ours(full_cfg, g)
    params = {"device": torch.device("cuda"), "nsp_weight": None}
    loss, lm_loss, nsp_loss, img_loss, nsp, lm = REF["train"].forward(enc, loader_batch(b), params, output_nsp_scores=True, output_lm_scores=True,
                                                                    evaluation=True)
    assert loss is None and lm_loss is None and nsp_loss is None and img_loss is None
    labels = b["mask"].to(nsp.device)
    a_, b_, c_ = lm.size()
    nll = F.cross_entropy(lm.view(a_ * b_, c_), labels.view(-1), ignore_index=-1, reduction="none").view(a_, b_)     # val_lm.py:124-136
    np.testing.assert_allclose((-nll.sum(-1)).cpu().numpy(), g["seq_score"], atol=1e-4, rtol=0)
    np.testing.assert_allclose(nsp.cpu().numpy(), g["nsp_scores"], atol=1e-4, rtol=0)


def test_train_forward_training_branch(full_cfg):
    """train.py:445-452: forward(encoder, batch, params, sample_size=None) -> loss, lm_loss, nsp_loss, img_loss."""
    g, b = load_golden("train6_perturbed")
    n = b["tokens"].shape[0]
    enc = ours(full_cfg, g)
    extra = {"next_sentence_labels": torch.from_numpy(g["next_sentence_label"]).unsqueeze(0),
             "image_target": torch.from_numpy(g["image_target"]).unsqueeze(0).expand(n, -1, -1).unsqueeze(0),
             "image_label": torch.from_numpy(g["image_label"]).unsqueeze(0).expand(n, -1).unsqueeze(0)}
    params = {"device": torch.device("cuda"), "nsp_weight": torch.from_numpy(g["nsp_weight"]), "lm_loss_coeff": 1.0, "nsp_loss_coeff": 1.0,
              "img_loss_coeff": 1.0}
    loss, lm_loss, nsp_loss, img_loss = REF["train"].forward(enc, loader_batch(b, extra), params, sample_size=None)
    print(f"train.forward over the drop-in: lm {lm_loss:.6f} nsp {nsp_loss:.6f} img {img_loss:.6f}")
    assert abs(lm_loss - g["lm_loss"].item()) < 1e-4 and abs(nsp_loss - g["nsp_loss"].item()) < 1e-4
    assert abs(img_loss - g["img_loss"].item()) < 1e-4
    assert abs(loss.item() - (g["lm_loss"].item() + g["nsp_loss"].item() + g["img_loss"].item())) < 3e-4


class _OnDevice(torch.nn.Module):
    """What nn.DataParallel's scatter does for the reference model: its callers hand over CPU tensors (train.py:113-129)."""

    def __init__(self, module):
        super().__init__()
        self.module = module.cuda()

    def forward(self, *args, **kw):
        mv = lambda x: x.cuda() if torch.is_tensor(x) else x
        return self.module(*[mv(a) for a in args], **{k: mv(v) for k, v in kw.items()})


def test_val_lm_evaluate_loop_same_metrics_as_the_reference_model(full_cfg):
    """val_lm.visdial_evaluate, unmodified, once over the reference's own model (eager fp32 PyTorch on this GPU) and once over
    the drop-in: identical ranks, identical R@k / mean / MRR / NDCG (north star: identical in fp32 mode)."""
    from oracle import encode_inputs as enc_o
    g, _ = load_golden("gen8_perturbed")
    rng = np.random.RandomState(99)
    n_img, n_rounds, n_opt = 2, 10, 20             # two images: NDCG.observe squeezes a batch of one away (visdial_metrics.py:145)
    imgs = [enc_o.synth_image(rng) for _ in range(n_img)]
    cols = [[] for _ in range(8)]
    draw = lambda k: rng.randint(1000, 30522, size=k).tolist()
    for _ in range(n_img):
        history = [draw(12)]
        for r in range(n_rounds):
            history = history + [draw(6)]                                          # the round's question
            for j in range(n_opt):
                out = REF["du"].encode_input_gen(history + [draw(int(rng.randint(1, 7)))], 1, enc_o.CLS, enc_o.SEP, enc_o.MASK, max_seq_len=256,
                                                 mask_prob=0, is_negtive=0)
                for c, o in zip(cols, out):
                    c.append(o)
            history = history + [draw(4)]                                          # its ground-truth answer joins the history
    tokens, segments, positions, sep_indices, labels, weights, att, co = (torch.cat(c, 0) for c in cols)
    shape = lambda t: t.view(n_img, n_rounds, n_opt, *t.shape[1:])
    batch = {"tokens": shape(tokens), "segments": shape(segments), "positions": shape(positions), "weights": shape(weights),
             "sep_indices": shape(sep_indices), "mask": shape(labels), "hist_len": torch.zeros(n_img, n_rounds, n_opt, dtype=torch.long),
             "txt_attention_mask": shape(att), "co_attention_mask": shape(co.unsqueeze(1).repeat(1, 37, 1)),
             "image_feat": torch.stack([i[0] for i in imgs]), "image_loc": torch.stack([i[1] for i in imgs]),
             "image_mask": torch.stack([i[2] for i in imgs]),
             "gt_option_inds": torch.from_numpy(rng.randint(0, n_opt, size=(n_img, n_rounds))), "round_id": torch.tensor([[4], [9]]),
             "gt_relevance": torch.from_numpy(rng.choice([0, 0, 0.5, 1.0], size=(n_img, n_opt)).astype(np.float32)),
             "image_id": torch.tensor([4242, 4243])}
    for k in ("gt_option_inds", "round_id", "gt_relevance"):          # the metric classes index them with device tensors
        batch[k] = batch[k].cuda()
    params = {"n_gpus": 0.2, "num_epochs": 1, "device": torch.device("cuda"), "nsp_weight": None}     # -> chunks of 25 sequences
    vl = REF["val_lm"]
    sd = golden_state_dict(full_cfg, g["weight_seed"], g["perturbed"])
    ref_model = _OnDevice(build_reference_encoder(REF, REF_ROOT, {"bert_pretrained." + k: v for k, v in sd.items()}))
    vl.ranks_json.clear()
    m_ref = vl.visdial_evaluate([batch], params, n_img, ref_model, ref_model)
    ranks_ref = [dict(r) for r in vl.ranks_json]
    del ref_model
    torch.cuda.empty_cache()
    enc = ours(full_cfg, g, max_sequences=25)
    vl.ranks_json.clear()
    m_ours = vl.visdial_evaluate([batch], params, n_img, enc, enc)
    ranks_ours = [dict(r) for r in vl.ranks_json]
    vl.ranks_json.clear()
    print("val_lm.visdial_evaluate: reference model", {k: round(float(v), 5) for k, v in m_ref.items() if "_round_" not in k})
    print("val_lm.visdial_evaluate: drop-in        ", {k: round(float(v), 5) for k, v in m_ours.items() if "_round_" not in k})
    assert ranks_ours == ranks_ref
    assert set(m_ref) == set(m_ours)
    for k in m_ref:
        assert float(m_ref[k]) == pytest.approx(float(m_ours[k]), abs=1e-9), k


def test_reference_training_loop_body_trains_through_the_drop_in(full_cfg):
    """train.py:445-463 with the reference's own pieces — ``train.forward`` (unmodified), ``GradScaler``, ``scaler.scale(loss).backward()``,
    ``scaler.step(optimizer)`` — over the drop-in after ``enable_training()``: the gradients autograd hands the module's parameters are the
    device backward's and match what the UNMODIFIED reference model got from the same ``loss.backward()`` (tests/golden/train6_grads.npz);
    a torch optimizer built from ``named_parameters()`` then updates the device masters in place and the next forward sees them."""
    import os
    g, b = load_golden("train6_perturbed")
    n = b["tokens"].shape[0]
    enc = ours(full_cfg, g).enable_training("fp16").eval()       # eval(): dropout off, as in the fixture; gradients still flow
    assert [k for k, _ in enc.named_parameters()] == ["bert_pretrained." + k for k in golden_state_dict(full_cfg, g["weight_seed"], g["perturbed"])
                                                      if k != "cls.predictions.decoder.weight"]
    extra = {"next_sentence_labels": torch.from_numpy(g["next_sentence_label"]).unsqueeze(0),
             "image_target": torch.from_numpy(g["image_target"]).unsqueeze(0).expand(n, -1, -1).unsqueeze(0),
             "image_label": torch.from_numpy(g["image_label"]).unsqueeze(0).expand(n, -1).unsqueeze(0)}
    params = {"device": torch.device("cuda"), "nsp_weight": torch.from_numpy(g["nsp_weight"]), "lm_loss_coeff": 1.0, "nsp_loss_coeff": 1.0,
              "img_loss_coeff": 1.0}
    optimizer = torch.optim.AdamW([p for p in enc.parameters() if p.requires_grad], lr=1e-4, weight_decay=0.01)
    scaler = torch.cuda.amp.GradScaler()
    loss, lm_loss, nsp_loss, img_loss = REF["train"].forward(enc, loader_batch(b, extra), params, sample_size=None)
    assert loss.requires_grad
    assert abs(lm_loss - g["lm_loss"].item()) < 2e-2 and abs(nsp_loss - g["nsp_loss"].item()) < 2e-2 and abs(img_loss - g["img_loss"].item()) < 2e-2
    scaler.scale(loss).backward()
    scale = scaler.get_scale()
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "train6_grads.npz"))
    named = dict(enc.named_parameters())
    gmax, worst, seen = float(z["grad_norm"].max()), 0.0, 0
    for name, norm, none in zip(z["names"], z["grad_norm"], z["grad_none"]):
        name = str(name)
        if name == "cls.predictions.decoder.weight":
            continue
        p = named["bert_pretrained." + name]
        if none:
            assert p.grad is None, name
            continue
        seen += 1
        worst = max(worst, abs(float(p.grad.double().norm()) / scale - norm) / max(norm, 1e-3 * gmax))
    print(f"reference loop over the drop-in: {seen} parameter gradients, L2 norms vs the reference model's own: worst relative difference {worst:.3e}")
    assert seen > 500 and worst < 5e-2
    before = named["bert_pretrained.bert.encoder.layer.3.output.dense.weight"].detach().clone()
    scaler.step(optimizer)
    scaler.update()
    optimizer.zero_grad()
    assert not torch.equal(before, named["bert_pretrained.bert.encoder.layer.3.output.dense.weight"].detach())
    loss2, *_ = REF["train"].forward(enc, loader_batch(b, extra), params, sample_size=None)
    print(f"loss {loss.item():.5f} -> {loss2.item():.5f} after one torch.optim.AdamW step through the reference's loop body")
    assert loss2.item() < loss.item()
    # the NSP scores are a differentiable output too (dense_annotation_finetuning.py:262-293 builds its losses on them in torch)
    out = enc(b["tokens"], b["image_feat"], b["image_loc"], sep_indices=b["sep_indices"], token_type_ids=b["segments"],
              token_position_ids=b["positions"], masked_lm_labels=b["mask"], attention_mask=b["txt_attention_mask"],
              next_sentence_label=torch.from_numpy(g["next_sentence_label"]).cuda(), output_nsp_scores=True,
              image_attention_mask=b["image_mask"], co_attention_mask=b["co_attention_mask"],
              image_label=torch.from_numpy(g["image_label"]).unsqueeze(0).expand(n, -1), image_target=torch.from_numpy(g["image_target"]).unsqueeze(0).expand(n, -1, -1),
              nsp_weight=None, lm_weight=b["weights"])
    scores = out[3]
    y = torch.from_numpy(g["next_sentence_label"]).cuda()
    torch.nn.functional.cross_entropy(scores, y).backward()
    want = (torch.softmax(scores.detach(), -1) - torch.nn.functional.one_hot(y, 2)).sum(0) / n
    got = named["bert_pretrained.cls.bi_seq_relationship.bias"].grad
    assert (got - want).abs().max().item() < 1e-5
    assert named["bert_pretrained.cls.imagePredictions.decoder.weight"].grad is None        # that loss was not part of this backward
    # train() mode: dropout on -> a different loss on the same batch
    enc.train()
    loss3, *_ = REF["train"].forward(enc, loader_batch(b, extra), params, sample_size=None)
    assert abs(loss3.item() - loss2.item()) > 1e-3
    loss3.backward()
    optimizer.zero_grad()
    # inference through the same module afterwards uses the UPDATED weights
    enc.eval()
    with torch.no_grad():
        o = enc(b["tokens"], b["image_feat"], b["image_loc"], token_type_ids=b["segments"], token_position_ids=b["positions"],
                masked_lm_labels=b["mask"], attention_mask=b["txt_attention_mask"], image_attention_mask=b["image_mask"],
                co_attention_mask=b["co_attention_mask"], output_nsp_scores=True)
    assert (o[3].cpu() - scores.detach().cpu()).abs().max().item() < 2e-2
