#!/usr/bin/env python
"""Generate the golden fixtures in this directory by running the UNMODIFIED reference.

Run in the build container only (needs the read-only reference checkout):

    python tests/golden/make_golden.py [--ref /root/reference] [--only case ...]

What it does
  1. Imports ``/root/reference/models/vilbert_dialog.py`` / ``visual_dialog_encoder.py`` /
     ``utils/data_utils.py`` / ``utils/visdial_metrics.py`` as they are.  Two third-party modules the
     reference imports but never uses on this path are absent offline and are stubbed
     (``pytorch_transformers``, ``pytorch_pretrained_bert``); ``Tensor.cuda`` is a no-op on CPU
     (the reference calls ``pe.cuda()`` on a buffer it never reads, models/vilbert_dialog.py:314).
  2. Builds ``VisualDialogEncoder`` without its network download (``__new__`` + a directly constructed
     ``BertForMultiModalPreTraining``) and loads ``unimm_b200.weights.random_state_dict`` through the
     reference's own ``load_state_dict(strict=True)`` — which also pins the 535-key layout.
  3. Builds inputs with the reference's own ``encode_input_gen/_dis/encode_input`` and
     ``encode_image_input`` and asserts that ``oracle.encode_inputs`` reproduces them bit for bit.
  4. Calls ``VisualDialogEncoder.forward`` with the keyword arguments ``train.forward`` uses
     (train.py:142-161), applies val_lm.py:131-136 to the full logits, and saves inputs + outputs.

Nothing here is imported by the product or by the GPU-side tests; the ``.npz`` files are.
"""
from __future__ import annotations

import argparse
import os
import random
import sys
import types

import numpy as np
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import encode_inputs as enc            # noqa: E402
from oracle import visdial_metrics as om           # noqa: E402
from unimm_b200.config import DEFAULT_CONFIG_PATH, ViLBertConfig   # noqa: E402
from unimm_b200.weights import PREFIX, random_state_dict            # noqa: E402


def import_reference(ref_root: str):
    for name in ("pytorch_transformers", "pytorch_transformers.modeling_bert", "pytorch_pretrained_bert",
                 "pytorch_pretrained_bert.file_utils"):
        sys.modules[name] = types.ModuleType(name)
    sys.modules["pytorch_transformers.modeling_bert"].BertEmbeddings = object
    sys.modules["pytorch_transformers"].modeling_bert = sys.modules["pytorch_transformers.modeling_bert"]
    sys.modules["pytorch_pretrained_bert.file_utils"].cached_path = lambda *a, **k: None
    torch.Tensor.cuda = lambda self, *a, **k: self
    sys.path.insert(0, ref_root)
    import importlib
    vd = importlib.import_module("models.vilbert_dialog")
    vde = importlib.import_module("models.visual_dialog_encoder")
    du = importlib.import_module("utils.data_utils")
    vm = importlib.import_module("utils.visdial_metrics")
    return vd, vde, du, vm


def build_reference_encoder(vd, vde, ref_root, sd):
    cfg = vd.BertConfig.from_json_file(os.path.join(ref_root, "config", "bert_base_6layer_6conect.json"))
    model = vde.VisualDialogEncoder.__new__(vde.VisualDialogEncoder)
    torch.nn.Module.__init__(model)
    model.bert_pretrained = vd.BertForMultiModalPreTraining(cfg)
    missing = model.load_state_dict(sd, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    assert len(model.state_dict()) == 535
    model.eval()
    return model


def ref_batch(du, context, answers, feats, loc, image_mask, fn, seed, **kw):
    """The reference's encoders driven like dataloader_visdial.py:394-455."""
    np.random.seed(seed)
    cols = [[] for _ in range(8)]
    for j, ans in enumerate(answers):
        args = dict(max_seq_len=256, mask_prob=0, is_negtive=0)
        args.update({k: (v[j] if isinstance(v, list) else v) for k, v in kw.items()})
        out = fn(context + [ans], 1, enc.CLS, enc.SEP, enc.MASK, **args)
        for c, o in zip(cols, out):
            c.append(o)
    tokens, segments, positions, sep_indices, labels, weights, att, co = (torch.cat(c, 0) for c in cols)
    n, R = tokens.shape[0], feats.shape[0]
    return {"tokens": tokens, "segments": segments, "positions": positions, "sep_indices": sep_indices,
            "mask": labels, "weights": weights, "txt_attention_mask": att,
            "co_attention_mask": co.unsqueeze(1).repeat(1, R, 1),
            "image_feat": feats.unsqueeze(0).expand(n, -1, -1).contiguous(),
            "image_loc": loc.unsqueeze(0).expand(n, -1, -1).contiguous(),
            "image_mask": image_mask.unsqueeze(0).expand(n, -1).contiguous()}


def oracle_batch(context, answers, feats, loc, image_mask, fn, seed, **kw):
    rng = np.random.RandomState(seed)
    cols = [[] for _ in range(8)]
    for j, ans in enumerate(answers):
        args = dict(max_seq_len=256, mask_prob=0, is_negative=0)
        for k, v in kw.items():
            k = "is_negative" if k == "is_negtive" else k
            args[k] = v[j] if isinstance(v, list) else v
        out = fn(context + [ans], 1, rng=rng, **args)
        for c, o in zip(cols, out):
            c.append(o)
    return [torch.cat(c, 0) for c in cols]


def assert_encoders_match(b, o):
    names = ["tokens", "segments", "positions", "sep_indices", "mask", "weights", "txt_attention_mask"]
    for n, t in zip(names, o[:7]):
        assert b[n].dtype == t.dtype or n == "txt_attention_mask", (n, b[n].dtype, t.dtype)
        assert torch.equal(b[n].long(), t.long()), f"oracle encoder mismatch in {n}"
    assert torch.equal(b["co_attention_mask"][:, 0, :], o[7]), "oracle encoder mismatch in co mask"


def call_reference(model, b, train_extras=None):
    """VisualDialogEncoder.forward with train.forward's keyword set (train.py:142-161)."""
    kw = dict(sep_indices=b["sep_indices"], sep_len=None, token_type_ids=b["segments"],
              token_position_ids=b["positions"], masked_lm_labels=b["mask"], attention_mask=b["txt_attention_mask"],
              next_sentence_label=None, output_nsp_scores=True, output_lm_scores=True,
              image_attention_mask=b["image_mask"], co_attention_mask=b["co_attention_mask"], image_label=None,
              image_target=None, nsp_weight=None, lm_weight=b["weights"])
    if train_extras:
        kw.update(train_extras)
    with torch.no_grad():
        return model(b["tokens"], b["image_feat"], b["image_loc"], **kw)


def val_lm_scores(lm_scores, labels):
    """val_lm.py:124-136."""
    a, bb, c = lm_scores.size()
    nll = F.cross_entropy(lm_scores.view(a * bb, c), labels.view(-1), ignore_index=-1, reduction="none").view(a, bb)
    return -nll.sum(-1), nll


def pack_inputs(b):
    out = {k: v.numpy() for k, v in b.items() if k not in ("image_feat", "image_loc", "image_mask", "co_attention_mask",
                                                           "txt_attention_mask")}
    out["txt_attention_mask"] = np.packbits(b["txt_attention_mask"].bool().numpy(), axis=-1)
    out["txt_attention_mask_is_long"] = np.array(b["txt_attention_mask"].dtype == torch.long)
    out["co_txt_mask"] = b["co_attention_mask"][:, 0, :].numpy().astype(np.uint8)
    return out


def save(name, **arrays):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **arrays)
    print(f"wrote {path}: {os.path.getsize(path) / 1024:.1f} KiB")


TAP_LAYERS = None


def add_taps(model, store):
    """Forward hooks on every encoder layer: keep a few rows of sequence 0 for bisecting."""
    enc_mod = model.bert_pretrained.bert.encoder
    hooks = []

    def mk(name, is_conn):
        def hook(_m, _inp, out):
            if is_conn:
                store[name + ".img"], store[name + ".txt"] = out[0][0].clone(), out[1][0].clone()
            else:
                store[name] = out[0][0].clone()
        return hook
    for i, l in enumerate(enc_mod.layer):
        hooks.append(l.register_forward_hook(mk(f"t{i}", False)))
    for i, l in enumerate(enc_mod.v_layer):
        hooks.append(l.register_forward_hook(mk(f"v{i}", False)))
    for i, l in enumerate(enc_mod.c_layer):
        hooks.append(l.register_forward_hook(mk(f"c{i}", True)))
    hooks.append(model.bert_pretrained.bert.embeddings.register_forward_hook(
        lambda _m, _i, out: store.__setitem__("emb.txt", out[0].clone())))
    hooks.append(model.bert_pretrained.bert.v_embeddings.register_forward_hook(
        lambda _m, _i, out: store.__setitem__("emb.img", out[0].clone())))
    return hooks


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    ap.add_argument("--only", nargs="*", default=None)
    args = ap.parse_args()
    want = lambda n: args.only is None or n in args.only

    vd, vde, du, vm = import_reference(args.ref)
    cfg = ViLBertConfig.from_json_file(DEFAULT_CONFIG_PATH)
    torch.set_num_threads(os.cpu_count())

    rng = np.random.RandomState(1234)
    context, answers = enc.synth_round(rng, n_candidates=100)
    feats, loc, image_mask = enc.synth_image(rng)
    # make sure the first 8 candidates cover the answer-length range 1..7 (+SEP → P in 2..8)
    for j, n in enumerate([1, 2, 3, 4, 5, 6, 7, 4]):
        answers[j] = rng.randint(1000, 30522, size=n).tolist()
    image_np = dict(image_feat=feats.numpy(), image_loc=loc.numpy(), image_mask=image_mask.numpy())

    models = {}

    def get_model(seed, perturbed):
        key = (seed, perturbed)
        if key not in models:
            models.clear()
            models[key] = build_reference_encoder(vd, vde, args.ref, random_state_dict(cfg, seed, perturbed, prefix=PREFIX))
        return models[key]

    # ---------------------------------------------------------------- generative, 8 candidates
    for name, seed, perturbed in (("gen8_default", 0, False), ("gen8_perturbed", 1, True)):
        if not want(name):
            continue
        b = ref_batch(du, context, answers[:8], feats, loc, image_mask, du.encode_input_gen, seed=7)
        assert_encoders_match(b, oracle_batch(context, answers[:8], feats, loc, image_mask, enc.encode_gen, seed=7))
        model = get_model(seed, perturbed)
        taps = {}
        hooks = add_taps(model, taps) if perturbed else []
        _, _, _, nsp, lm = call_reference(model, b)
        for h in hooks:
            h.remove()
        score, nll = val_lm_scores(lm, b["mask"])
        rows = (b["mask"] != -1).nonzero()
        logits_rows = lm[rows[:, 0], rows[:, 1]]
        lab = b["mask"][rows[:, 0], rows[:, 1]]
        ul = torch.log(torch.clamp(1.0 - F.softmax(logits_rows, -1), min=1e-6)).gather(1, lab[:, None])[:, 0]
        extra = {}
        if taps:
            # rows of sequence 0: CLS, first/last context, first/last A, first/last B, first pad
            L = int((b["txt_attention_mask"][0, 0].sum() + b["co_attention_mask"][0, 0].sum() + 1) // 2)
            T = int(b["txt_attention_mask"][0, 0].sum())
            ctx = 2 * L - T
            trow = [0, 1, ctx - 1, ctx, L - 1, L, T - 1, min(T, 255)]
            extra["tap_txt_rows"] = np.array(trow)
            extra["tap_img_rows"] = np.array([0, 17, 36])
            for k, v in taps.items():
                is_img = k.endswith(".img") or k.startswith("v")
                extra["tap." + k] = v[[0, 17, 36]].numpy() if is_img else v[trow].numpy()
        save(name, weight_seed=np.array(seed), perturbed=np.array(perturbed), **pack_inputs(b), **image_np,
             seq_score=score.numpy(), nsp_scores=nsp.numpy(), token_rows=rows.numpy(),
             token_logp=(-nll[rows[:, 0], rows[:, 1]]).numpy(), token_ul=ul.numpy(),
             logits_row0_first64=lm[0, rows[0, 1], :64].numpy(), **extra)

    # ---------------------------------------------------------------- discriminative, 8 candidates (val.py path)
    if want("dis8_perturbed"):
        b = ref_batch(du, context, answers[:8], feats, loc, image_mask, du.encode_input_dis, seed=11,
                      mask_prob=0.15, vocab_size=30522)
        assert_encoders_match(b, oracle_batch(context, answers[:8], feats, loc, image_mask, enc.encode_dis, seed=11,
                                              mask_prob=0.15, vocab_size=30522))
        model = get_model(1, True)
        _, _, _, nsp, lm = call_reference(model, b)
        score, nll = val_lm_scores(lm, b["mask"])
        rows = (b["mask"] != -1).nonzero()
        save("dis8_perturbed", weight_seed=np.array(1), perturbed=np.array(True), **pack_inputs(b), **image_np,
             seq_score=score.numpy(), nsp_scores=nsp.numpy(), nsp_prob0=F.softmax(nsp, 1)[:, 0].numpy(),
             token_rows=rows.numpy(), token_logp=(-nll[rows[:, 0], rows[:, 1]]).numpy())

    # ---------------------------------------------------------------- train forward + losses (config 3 shape, B=6)
    if want("train6_perturbed"):
        np.random.seed(23)
        random.seed(23)
        neg = [0, 1, 1, 0, 1, 1]
        cols = [[] for _ in range(8)]
        for j in range(6):
            out = du.encode_input(0.5, context + [answers[j]], 1, enc.CLS, enc.SEP, enc.MASK, max_seq_len=256,
                                  mask_prob=0.15, is_negtive=neg[j], weight=1, vocab_size=30522)
            for c, o in zip(cols, out):
                c.append(o.long() if o.dim() == 3 else o)
        o_rng = np.random.RandomState(23)
        ocols = [[] for _ in range(8)]
        for j in range(6):
            out = enc.encode(0.5, context + [answers[j]], 1, rng=o_rng, max_seq_len=256, mask_prob=0.15,
                             is_negative=neg[j], weight=1, vocab_size=30522)
            for c, o in zip(ocols, out):
                c.append(o.long() if o.dim() == 3 else o)
        tokens, segments, positions, sep_indices, labels, weights, att, co = (torch.cat(c, 0) for c in cols)
        for a_, b_ in zip((tokens, segments, positions, sep_indices, labels, weights, att, co),
                          (torch.cat(c, 0) for c in ocols)):
            assert torch.equal(a_, b_), "oracle encode() mismatch"
        target = torch.from_numpy(np.random.dirichlet(np.ones(1601), size=37).astype(np.float32))
        f2, l2, im2, tgt2, img_label = du.encode_image_input(feats.numpy(), 37, loc.numpy(), target.numpy(),
                                                             max_regions=37, mask_prob=0.15)
        b = {"tokens": tokens, "segments": segments, "positions": positions, "sep_indices": sep_indices, "mask": labels,
             "weights": weights, "txt_attention_mask": att, "co_attention_mask": co.unsqueeze(1).repeat(1, 37, 1),
             "image_feat": f2.unsqueeze(0).expand(6, -1, -1).contiguous(),
             "image_loc": l2.unsqueeze(0).expand(6, -1, -1).contiguous(),
             "image_mask": im2.unsqueeze(0).expand(6, -1).contiguous()}
        nsl = torch.LongTensor(neg)
        extras = dict(next_sentence_label=nsl, image_label=img_label.unsqueeze(0).expand(6, -1).contiguous(),
                      image_target=tgt2.unsqueeze(0).expand(6, -1, -1).contiguous(),
                      nsp_weight=torch.FloatTensor([[5.0, 1.0]]))
        model = get_model(1, True)
        lm_loss, img_loss, nsp_loss, nsp, lm = call_reference(model, b, extras)
        save("train6_perturbed", weight_seed=np.array(1), perturbed=np.array(True), **pack_inputs(b),
             image_feat=f2.numpy(), image_loc=l2.numpy(), image_mask=im2.numpy(),
             next_sentence_label=nsl.numpy(), image_label=img_label.numpy(), image_target=tgt2.numpy(),
             nsp_weight=np.array([[5.0, 1.0]], dtype=np.float32),
             lm_loss=lm_loss.numpy(), img_loss=img_loss.numpy(), nsp_loss=nsp_loss.numpy(), nsp_scores=nsp.numpy())

    # ---------------------------------------------------------------- backward of the train step: the reference's own
    # autograd on the train6_perturbed batch — loss = lm + nsp + img as train.py:163-168 builds it, loss.backward() as :453 — with the
    # dropout layers in eval mode.  250 M gradient values do not fit a fixture: per parameter its L2 norm and sum, a few small tensors in full.
    if want("train6_grads"):
        sys.path.insert(0, os.path.dirname(HERE))
        from conftest import load_golden
        g6, b6 = load_golden("train6_perturbed")
        extras = dict(next_sentence_label=torch.from_numpy(g6["next_sentence_label"]),
                      image_label=torch.from_numpy(g6["image_label"]).unsqueeze(0).expand(6, -1).contiguous(),
                      image_target=torch.from_numpy(g6["image_target"]).unsqueeze(0).expand(6, -1, -1).contiguous(),
                      nsp_weight=torch.from_numpy(g6["nsp_weight"]))
        model = get_model(int(g6["weight_seed"]), bool(g6["perturbed"]))
        model.eval()
        for prm in model.parameters():
            prm.grad = None
        kw = dict(sep_indices=b6["sep_indices"], sep_len=None, token_type_ids=b6["segments"], token_position_ids=b6["positions"],
                  masked_lm_labels=b6["mask"], attention_mask=b6["txt_attention_mask"], output_nsp_scores=False, output_lm_scores=False,
                  image_attention_mask=b6["image_mask"], co_attention_mask=b6["co_attention_mask"], lm_weight=b6["weights"], **extras)
        lm_loss, img_loss, nsp_loss = model(b6["tokens"], b6["image_feat"], b6["image_loc"], **kw)
        loss = 1.0 * lm_loss.mean() + 1.0 * nsp_loss.mean() + 1.0 * img_loss.mean()                 # train.py:163-168
        loss.backward()                                                                             # train.py:453 (without the GradScaler factor)
        names, norms, sums, none = [], [], [], []
        full = {}
        keep = ("bert.encoder.layer.11.output.LayerNorm.weight", "bert.encoder.layer.0.attention.self.query.bias",
                "bert.encoder.c_layer.2.biOutput.LayerNorm1.bias", "bert.encoder.v_layer.3.attention.output.dense.bias",
                "cls.bi_seq_relationship.weight", "cls.predictions.transform.dense.bias", "bert.embeddings.token_type_embeddings_extension.weight",
                "bert.v_embeddings.image_location_embeddings.weight", "bert.t_pooler.dense.bias")
        for n_, prm in model.named_parameters():
            k = n_[len("bert_pretrained."):] if n_.startswith("bert_pretrained.") else n_
            names.append(k)
            if prm.grad is None:
                none.append(True), norms.append(0.0), sums.append(0.0)
                continue
            none.append(False), norms.append(float(prm.grad.double().norm())), sums.append(float(prm.grad.double().sum()))
            if k in keep:
                full["grad__" + k] = prm.grad.numpy().copy()
        assert len(full) == len(keep), sorted(set(keep) - {k[6:] for k in full})
        save("train6_grads", names=np.array(names), grad_norm=np.array(norms), grad_sum=np.array(sums), grad_none=np.array(none),
             loss=np.float32(loss.detach()), **full)

    # ---------------------------------------------------------------- config 5: dense-annotation fine-tuning forward + loss
    # (dense_annotation_finetuning.py:253 -> train.forward; dataloader_dense_annotations.py:148-172: ONE mode per image,
    # relevance as the token weight -> the LongTensor truncates 0.2..0.8 to 0, relevance 0 makes the option a negative;
    # nsp_weight = None, dense_annotation_finetuning.py:145)
    for name, fn_name in (("ft8gen_perturbed", "encode_input_gen"), ("ft8dis_perturbed", "encode_input_dis")):
        if not want(name):
            continue
        np.random.seed(31)
        random.seed(31)
        relevance = [1.0, 0.0, 0.4, 0.8, 0.0, 1.0, 0.6, 1.0]
        kw = dict(mask_prob=0.1, vocab_size=30522, is_negtive=[int(r == 0) for r in relevance],
                  weight=[(r if r > 0 else 1) for r in relevance])
        b = ref_batch(du, context, answers[:8], feats, loc, image_mask, getattr(du, fn_name), seed=31, **kw)
        ofn = enc.encode_gen if fn_name.endswith("gen") else enc.encode_dis
        assert_encoders_match(b, oracle_batch(context, answers[:8], feats, loc, image_mask, ofn, seed=31, **kw))
        np.random.seed(32)
        target = torch.from_numpy(np.random.dirichlet(np.ones(1601), size=37).astype(np.float32))
        f2, l2, im2, tgt2, img_label = du.encode_image_input(feats.numpy(), 37, loc.numpy(), target.numpy(),
                                                             max_regions=37, mask_prob=0.1)
        b["image_feat"] = f2.unsqueeze(0).expand(8, -1, -1).contiguous()
        b["image_loc"] = l2.unsqueeze(0).expand(8, -1, -1).contiguous()
        b["image_mask"] = im2.unsqueeze(0).expand(8, -1).contiguous()
        nsl = torch.LongTensor([int(r == 0) for r in relevance])
        extras = dict(next_sentence_label=nsl, image_label=img_label.unsqueeze(0).expand(8, -1).contiguous(),
                      image_target=tgt2.unsqueeze(0).expand(8, -1, -1).contiguous(), nsp_weight=None)
        model = get_model(1, True)
        lm_loss, img_loss, nsp_loss, nsp, lm = call_reference(model, b, extras)
        print(name, "weights used:", sorted(set(b["weights"].view(-1).tolist())), "losses", lm_loss.item(), img_loss.item(), nsp_loss.item())
        save(name, weight_seed=np.array(1), perturbed=np.array(True), **pack_inputs(b),
             image_feat=f2.numpy(), image_loc=l2.numpy(), image_mask=im2.numpy(),
             next_sentence_label=nsl.numpy(), image_label=img_label.numpy(), image_target=tgt2.numpy(),
             relevance=np.array(relevance, dtype=np.float32),
             lm_loss=lm_loss.numpy(), img_loss=img_loss.numpy(), nsp_loss=nsp_loss.numpy(), nsp_scores=nsp.numpy())

    # ---------------------------------------------------------------- dense-annotation objective on NSP probabilities
    if want("rankloss"):
        import importlib
        from oracle import rank_loss as orl
        rl = importlib.import_module("utils.rank_loss")
        g = np.random.RandomState(77)
        cases = []
        for kind in range(5):
            b = 3
            p = g.rand(b, 100).astype(np.float32)
            if kind == 1:
                p = (g.rand(b, 100) ** 4).astype(np.float32)            # peaked, NSP-like
            if kind == 3:
                p = np.sort(p, -1)[:, ::-1].copy()                       # already sorted
            y = g.choice([0, 0, 0, 0.2, 0.4, 0.6, 0.8, 1.0], size=(b, 100)).astype(np.float32)
            if kind == 2:
                y[1] = 0                                                 # a slate without any relevant option
            if kind == 4:
                y[:] = 0                                                 # none at all: the loss is 0
            loss = rl.neuralNDCG_transposed(torch.from_numpy(p.copy()), torch.from_numpy(y.copy()))
            mine = orl.neural_ndcg_transposed(p, y)[0]
            assert abs(float(loss) - float(mine)) < 1e-6, (kind, float(loss), float(mine))
            cases.append((p, y, np.float32(loss)))
        # val.py:152-161 (inline code in the reference; the same torch expressions)
        probs = torch.from_numpy(g.rand(5, 4, 10, 100).astype(np.float32))
        res = None
        for tmp in probs:
            a_ = tmp.min(dim=-1)[0].unsqueeze(-1)
            b_ = tmp.max(dim=-1)[0].unsqueeze(-1)
            e_x = (tmp - a_) / (b_ - a_)
            res_tmp = e_x / (e_x.sum(dim=-1).unsqueeze(-1))
            res = res_tmp if res is None else res + res_tmp
        assert np.abs(orl.ensemble_normalise(probs.numpy().reshape(5, 40, 100)).reshape(4, 10, 100) - res.numpy()).max() < 1e-6
        save("rankloss", y_pred=np.stack([c[0] for c in cases]), y_true=np.stack([c[1] for c in cases]),
             loss=np.array([c[2] for c in cases], dtype=np.float32), ens_probs=probs.numpy(), ens_out=res.numpy())


    # ---------------------------------------------------------------- gradient of the dense-annotation objective
    if want("rankloss_grad"):
        import importlib
        from oracle import rank_loss as orl
        rl = importlib.import_module("utils.rank_loss")
        g = np.random.RandomState(78)
        ps, ys, grads, losses = [], [], [], []
        for kind in range(4):
            p = g.rand(3, 100).astype(np.float32)
            if kind == 1:
                p = (g.rand(3, 100) ** 4).astype(np.float32)            # peaked, NSP-like
            y = g.choice([0, 0, 0, 0.2, 0.4, 0.6, 0.8, 1.0], size=(3, 100)).astype(np.float32)
            if kind == 2:
                y[1] = 0                                                 # a slate without any relevant option: no gradient there
            if kind == 3:
                p[0, 10:14] = p[0, 10]                                   # tied scores (abs'(0) = 0)
            tp = torch.from_numpy(p.copy()).requires_grad_()
            loss = rl.neuralNDCG_transposed(tp, torch.from_numpy(y.copy()))
            loss.backward()                                             # dense_annotation_finetuning.py:296
            mine = orl.neural_ndcg_transposed_grad(p, y)
            assert np.abs(mine - tp.grad.numpy()).max() < 1e-5 * np.abs(tp.grad.numpy()).max(), kind
            ps.append(p), ys.append(y), grads.append(tp.grad.numpy().copy()), losses.append(np.float32(loss.detach()))
        save("rankloss_grad", y_pred=np.stack(ps), y_true=np.stack(ys), grad=np.stack(grads), loss=np.array(losses, dtype=np.float32))

    # ---------------------------------------------------------------- bench-shape sweep slice: 3 rounds x 100 candidates of ONE image
    # (val_lm.py:104-137 over several rounds of an image: contexts of different lengths share one feature block).  Inputs are the
    # bench's own synthetic generator (unimm_b200.synthetic.synth_dialog_rounds), rebuilt here with the REFERENCE encoders and
    # required to be identical, so that the fixture pins exactly what bench.py / val_sweep.py feed the packed path.
    for name, seed, perturbed in (("sweep3x100_default", 0, False), ("sweep3x100_perturbed", 1, True)):
        if not want(name):
            continue
        from unimm_b200 import synthetic as syn
        image_id, round_ids = 7, (1, 5, 10)
        srng = np.random.RandomState(100003 + image_id)                      # same draw order as syn.synth_dialog_rounds
        s_feat, s_loc, s_mask = (torch.from_numpy(a) for a in syn.synth_image(srng))
        batches = []
        for r in round_ids:
            ctx_utts, ans = syn.synth_context(srng, r), syn.synth_answers(srng, 100)
            batches.append(ref_batch(du, ctx_utts, ans, s_feat, s_loc, s_mask, du.encode_input_gen, seed=7))
        (feat2, loc2, mask2), views = syn.synth_dialog_rounds(image_id, rounds=round_ids)
        assert np.array_equal(feat2, s_feat.numpy()) and np.array_equal(loc2, s_loc.numpy())
        for b, v in zip(batches, views):
            assert np.array_equal(b["tokens"].numpy(), v.tokens) and np.array_equal(b["segments"].numpy(), v.segments)
            assert np.array_equal(b["positions"].numpy(), v.positions) and np.array_equal(b["mask"].numpy(), v.labels)
            from unimm_b200.descriptors import dense_co_mask, dense_text_mask
            d = torch.from_numpy(v.desc)
            assert torch.equal(dense_text_mask(d, 256), b["txt_attention_mask"].bool())
            assert torch.equal(dense_co_mask(d, 256), b["co_attention_mask"][:, 0, :])
        model = get_model(seed, perturbed)
        scores = []
        for b in batches:
            for s in range(0, 100, 25):
                bb = {k: v[s:s + 25] for k, v in b.items()}
                _, _, _, nsp, lm = call_reference(model, bb)
                scores.append(val_lm_scores(lm, bb["mask"])[0])
        score = torch.cat(scores).view(len(round_ids), 100)
        ranks = vm.scores_to_ranks(score.view(1, len(round_ids), 100).clone()).view(len(round_ids), 100)
        save(name, weight_seed=np.array(seed), perturbed=np.array(perturbed), image_id=np.array(image_id), round_ids=np.array(round_ids),
             tokens=np.concatenate([v.tokens for v in views]).astype(np.int32), labels=np.concatenate([v.labels for v in views]).astype(np.int32),
             desc=np.concatenate([v.desc for v in views]), seq_score=score.numpy(), ranks=ranks.numpy())

    # ---------------------------------------------------------------- sequences truncated at max_seq_len (data_utils.py:205-209, :237-244)
    # context of 247 positions: answers of 1..3 tokens fit (T <= 256), 4..7 lose part of the masked copy, 8 loses all of it (L = 256),
    # 10 loses part of the visible copy too (L > 256)
    if want("gen10_truncated"):
        trng = np.random.RandomState(4321)
        draw = lambda n: trng.randint(1000, 30522, size=n).tolist()
        t_context = [draw(20)] + [draw(10) for _ in range(19)] + [draw(15)]          # 1 + 21 + 20 + 190 + 15 = 247 positions
        t_answers = [draw(n) for n in (1, 3, 4, 5, 7, 8, 10, 2, 6, 3)]
        b = ref_batch(du, t_context, t_answers, feats, loc, image_mask, du.encode_input_gen, seed=7)
        assert_encoders_match(b, oracle_batch(t_context, t_answers, feats, loc, image_mask, enc.encode_gen, seed=7))
        assert int(b["txt_attention_mask"][0, 0].sum()) == 251 and int(b["txt_attention_mask"][2, 0].sum()) == 256
        model = get_model(1, True)
        _, _, _, nsp, lm = call_reference(model, b)
        score, nll = val_lm_scores(lm, b["mask"])
        rows = (b["mask"] != -1).nonzero()
        save("gen10_truncated", weight_seed=np.array(1), perturbed=np.array(True), **pack_inputs(b), **image_np,
             answer_lens=np.array([len(a) for a in t_answers]), seq_score=score.numpy(), nsp_scores=nsp.numpy(),
             token_rows=rows.numpy(), token_logp=(-nll[rows[:, 0], rows[:, 1]]).numpy())

    # ---------------------------------------------------------------- config 4: 100 candidates, discriminative NSP ranking (val.py:125-161)
    if want("dis100_default"):
        b = ref_batch(du, context, answers, feats, loc, image_mask, du.encode_input_dis, seed=13)
        model = get_model(0, False)
        nsps = []
        for s in range(0, 100, 25):
            bb = {k: v[s:s + 25] for k, v in b.items()}
            _, _, _, nsp, lm = call_reference(model, bb)
            nsps.append(nsp)
        nsp = torch.cat(nsps)
        prob0 = F.softmax(nsp, 1)[:, 0]                                   # val.py:127-131: the "is the right answer" probability
        ranks = vm.scores_to_ranks(prob0.view(1, 1, 100).clone())
        srt = prob0.sort(descending=True)[0]
        print("dis100_default min adjacent gap", float((srt[:-1] - srt[1:]).min()))
        save("dis100_default", weight_seed=np.array(0), perturbed=np.array(False), **pack_inputs(b), **image_np,
             nsp_scores=nsp.numpy(), nsp_prob0=prob0.numpy(), ranks=ranks.view(100).numpy())


    # ---------------------------------------------------------------- config 3 at its stated size: train.py step forward + loss,
    # batch 240 = 40 images x 6 sequences (1 positive + 5 negatives of one round, dataloader_visdial.py:199-262), mode drawn per
    # sequence (train_dis_rate 0.5), mask_prob 0.15, unlikelihood weight -1 on the negatives, nsp_weight [5, 1] (train.py:402).
    # Image blocks are regenerated from seeds at test time (oracle.encode_inputs.synth_image / a seeded Dirichlet); only the
    # random masking decisions of encode_image_input are stored.
    if want("train240_perturbed"):
        np.random.seed(101)
        random.seed(101)
        n_img, per_img = 40, 6
        cols = [[] for _ in range(8)]
        nsl, seq_image, img_seeds = [], [], []
        feats_all, loc_all, mask_all, tgt_all, lab_all, zeroed = [], [], [], [], [], []
        for i in range(n_img):
            irng = np.random.RandomState(7000 + i)
            img_seeds.append(7000 + i)
            f_i, l_i, m_i = enc.synth_image(irng)
            tgt = torch.from_numpy(np.random.RandomState(8000 + i).dirichlet(np.ones(1601), size=37).astype(np.float32))
            f2, l2, im2, tgt2, img_label = du.encode_image_input(f_i.numpy(), 37, l_i.numpy(), tgt.numpy(), max_regions=37, mask_prob=0.15)
            feats_all.append(f2), loc_all.append(l2), mask_all.append(im2), tgt_all.append(tgt2), lab_all.append(img_label)
            zeroed.append((f2.abs().sum(-1) == 0).numpy())
            rnd = int(irng.randint(1, 11))
            ctx_i = [irng.randint(1000, 30522, size=int(irng.randint(5, 20))).tolist()]
            for _ in range(2 * (rnd - 1)):
                ctx_i.append(irng.randint(1000, 30522, size=int(irng.randint(2, 12))).tolist())
            ctx_i.append(irng.randint(1000, 30522, size=int(irng.randint(3, 10))).tolist())          # the round's question
            for j in range(per_img):
                neg = int(j > 0)
                ans = irng.randint(1000, 30522, size=int(irng.randint(1, 8))).tolist()
                out = du.encode_input(0.5, ctx_i + [ans], 1, enc.CLS, enc.SEP, enc.MASK, max_seq_len=256, mask_prob=0.15, is_negtive=neg,
                                      weight=1, vocab_size=30522)
                for c, o in zip(cols, out):
                    c.append(o.long() if o.dim() == 3 else o)
                nsl.append(neg), seq_image.append(i)
        tokens, segments, positions, sep_indices, labels, weights, att, co = (torch.cat(c, 0) for c in cols)
        seq_image = torch.LongTensor(seq_image)
        B = tokens.shape[0]
        stack = lambda xs: torch.stack(xs)[seq_image]
        b = {"tokens": tokens, "segments": segments, "positions": positions, "sep_indices": sep_indices, "mask": labels, "weights": weights,
             "txt_attention_mask": att, "co_attention_mask": co.unsqueeze(1).repeat(1, 37, 1),
             "image_feat": stack(feats_all), "image_loc": stack(loc_all), "image_mask": stack(mask_all)}
        extras = dict(next_sentence_label=torch.LongTensor(nsl), image_label=stack(lab_all), image_target=stack(tgt_all),
                      nsp_weight=torch.FloatTensor([[5.0, 1.0]]))
        model = get_model(1, True)
        lm_loss, img_loss, nsp_loss, nsp, lm = call_reference(model, b, extras)
        del lm
        print("train240: modes dis/gen", int((co[:, 0] == 1).sum()), int((co[:, 0] == 0).sum()), "losses", lm_loss.item(), img_loss.item(), nsp_loss.item())
        inp = pack_inputs(b)
        save("train240_perturbed", weight_seed=np.array(1), perturbed=np.array(True), **inp, seq_image=seq_image.numpy(),
             image_seeds=np.array(img_seeds), image_zeroed=np.stack(zeroed), image_label=torch.stack(lab_all).numpy(),
             image_loc=torch.stack(loc_all).numpy(), image_mask=torch.stack(mask_all).numpy(),
             next_sentence_label=np.array(nsl), nsp_weight=np.array([[5.0, 1.0]], dtype=np.float32),
             lm_loss=lm_loss.numpy(), img_loss=img_loss.numpy(), nsp_loss=nsp_loss.numpy(), nsp_scores=nsp.numpy())

    # ---------------------------------------------------------------- config 5 at its stated size: dense-annotation fine-tuning step,
    # the 100 options of one annotated round, ONE mode for all of them, relevance as token weight (integer-truncated),
    # nsp_weight None, + the objectives dense_annotation_finetuning.py:263-296 builds on the NSP scores
    for name, fn_name in (("ft100gen_perturbed", "encode_input_gen"), ("ft100dis_perturbed", "encode_input_dis")):
        if not want(name):
            continue
        import importlib
        rl = importlib.import_module("utils.rank_loss")
        np.random.seed(41)
        random.seed(41)
        relevance = np.random.RandomState(42).choice([0, 0, 0, 0.2, 0.4, 0.6, 0.8, 1.0], size=100).astype(np.float32)
        relevance[0] = 1.0
        rel_l = relevance.tolist()
        kw = dict(mask_prob=0.1, vocab_size=30522, is_negtive=[int(r == 0) for r in rel_l], weight=[(r if r > 0 else 1) for r in rel_l])
        b = ref_batch(du, context, answers, feats, loc, image_mask, getattr(du, fn_name), seed=41, **kw)
        np.random.seed(43)
        target = torch.from_numpy(np.random.RandomState(44).dirichlet(np.ones(1601), size=37).astype(np.float32))
        f2, l2, im2, tgt2, img_label = du.encode_image_input(feats.numpy(), 37, loc.numpy(), target.numpy(), max_regions=37, mask_prob=0.1)
        b["image_feat"] = f2.unsqueeze(0).expand(100, -1, -1).contiguous()
        b["image_loc"] = l2.unsqueeze(0).expand(100, -1, -1).contiguous()
        b["image_mask"] = im2.unsqueeze(0).expand(100, -1).contiguous()
        nsl = torch.LongTensor([int(r == 0) for r in rel_l])
        extras = dict(next_sentence_label=nsl, image_label=img_label.unsqueeze(0).expand(100, -1).contiguous(),
                      image_target=tgt2.unsqueeze(0).expand(100, -1, -1).contiguous(), nsp_weight=None)
        model = get_model(1, True)
        lm_loss, img_loss, nsp_loss, nsp, lm = call_reference(model, b, extras)
        del lm
        # dense_annotation_finetuning.py:263-296
        nsp_scores = nsp.view(-1, 100, 2)
        ce_nsp = F.cross_entropy(nsp_scores.view(-1, 2), nsl.view(-1))
        gt_rel = torch.from_numpy(relevance).view(1, 100)
        nsp_probs = F.softmax(nsp_scores, dim=-1)
        target_loss = rl.neuralNDCG_transposed(nsp_probs[:, :, 0], gt_rel)
        total = target_loss + lm_loss.mean() + 1.0 * ce_nsp
        print(name, "losses", lm_loss.item(), img_loss.item(), nsp_loss.item(), "neuralNDCG", float(target_loss), "total", float(total))
        save(name, weight_seed=np.array(1), perturbed=np.array(True), **pack_inputs(b),
             image_feat=f2.numpy(), image_loc=l2.numpy(), image_mask=im2.numpy(),
             next_sentence_label=nsl.numpy(), image_label=img_label.numpy(), image_target=tgt2.numpy(), relevance=relevance,
             lm_loss=lm_loss.numpy(), img_loss=img_loss.numpy(), nsp_loss=nsp_loss.numpy(), nsp_scores=nsp.numpy(),
             nsp_ce_unweighted=ce_nsp.numpy(), neural_ndcg_loss=np.float32(target_loss), total_loss=np.float32(total))

    # ---------------------------------------------------------------- config 1: 100 candidates, ranking metrics
    for name, seed, perturbed in (("gen100_default", 0, False),):
        if not want(name):
            continue
        b = ref_batch(du, context, answers, feats, loc, image_mask, du.encode_input_gen, seed=7)
        model = get_model(seed, perturbed)
        scores, nsps = [], []
        for s in range(0, 100, 25):                                  # chunks of 25
            bb = {k: v[s:s + 25] for k, v in b.items()}
            _, _, _, nsp, lm = call_reference(model, bb)
            scores.append(val_lm_scores(lm, bb["mask"])[0])
            nsps.append(nsp)
        score = torch.cat(scores)
        ranks = vm.scores_to_ranks(score.view(1, 1, 100).clone())
        sm = vm.SparseGTMetrics()
        sm.observe(score.view(1, 1, 100), torch.zeros(1, 1, dtype=torch.long))   # gt option is index 0
        rel = torch.from_numpy(np.random.RandomState(5).choice([0, 0, 0, 0.2, 0.4, 0.6, 0.8, 1.0], size=100)
                               .astype(np.float32)).view(1, 100)
        rel[0, 0] = 1.0
        nd = vm.NDCG()
        # NDCG.observe squeezes a batch of one away (visdial_metrics.py:145); val_lm feeds 2 images per batch
        nd.observe(score.view(1, 100).repeat(2, 1), rel.repeat(2, 1))
        metrics = {**{k: v for k, v in sm.retrieve().items() if "_round_" not in k}, **nd.retrieve()}
        assert torch.equal(om.scores_to_ranks(score.view(1, 1, 100)), ranks)
        mine = {**om.sparse_metrics(score.view(1, 1, 100), torch.zeros(1, 1, dtype=torch.long)),
                "ndcg": om.ndcg(score.view(1, 100), rel)}
        for k in metrics:
            assert abs(metrics[k] - mine[k]) < 1e-6, (k, metrics[k], mine[k])
        srt = score.sort(descending=True)[0]
        print(name, "metrics", metrics, "min adjacent gap", float((srt[:-1] - srt[1:]).min()))
        save(name, weight_seed=np.array(seed), perturbed=np.array(perturbed), **pack_inputs(b), **image_np,
             seq_score=score.numpy(), nsp_scores=torch.cat(nsps).numpy(), ranks=ranks.view(100).numpy(),
             relevance=rel.numpy(), metric_names=np.array(sorted(metrics)),
             metric_values=np.array([metrics[k] for k in sorted(metrics)], dtype=np.float64))


if __name__ == "__main__":
    main()
