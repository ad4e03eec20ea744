"""ORACLE (test infrastructure, not product code): CPU restatement of the reference hot path.

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` may import this module; the product (``unimm_b200/``) never does.

What it restates (all citations are ``/root/reference`` paths): the two-stream ViLBERT forward of
``BertForMultiModalPreTraining`` (models/vilbert_dialog.py:1519-1626) as plain functional
``torch`` on CPU, reading weights from a state dict in the reference's own key layout, plus the
val_lm scoring rule (val_lm.py:131-137).  Every arithmetic op on this path is a ``torch`` op in the
reference too (no third-party arithmetic), so the restatement calls the same ops in
the same order; it differs only in being a set of functions instead of an ``nn.Module`` tree.

Parity pinning: the reference ships no tests or golden vectors.  The oracle is pinned
against the *reference itself* executed in the build container: ``tests/golden/make_golden.py``
imports ``/root/reference/models/vilbert_dialog.py`` unmodified, loads the weights of
``unimm_b200.weights.random_state_dict`` through the reference's own ``load_state_dict``, feeds inputs
built by the reference's own ``utils/data_utils.encode_input_*`` and commits the outputs as
``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` checks this file against them.

Dropout: every ``nn.Dropout`` on the path is the identity here (the reference's eval mode;
SURVEY.md §7 "Dropout") unless the caller passes ``drop``: a callable ``(site, tensor) -> tensor`` invoked at exactly the
reference's ``nn.Dropout`` call sites (models/vilbert_dialog.py:355, :405, :424, :467, :534, :553, :596, :693, :716, :746, :749,
:1065, :1491) with a stable site name — the training-step tests use it to apply the SAME keep-masks the device step draws, so
that the step with dropout on can be checked against ``torch.autograd`` exactly.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch
import torch.nn.functional as F

MASK_NEG = -10000.0  # models/vilbert_dialog.py:1418,1423,1431


def _w(sd, name, dtype):
    return sd[name].to(dtype)


def linear(sd, name: str, x: torch.Tensor) -> torch.Tensor:
    """``nn.Linear`` named ``name`` (weight [out,in], bias [out])."""
    return F.linear(x, _w(sd, name + ".weight", x.dtype), _w(sd, name + ".bias", x.dtype))


def layer_norm(sd, name: str, x: torch.Tensor) -> torch.Tensor:
    """``BertLayerNorm = torch.nn.LayerNorm(eps=1e-12)`` (models/vilbert_dialog.py:279, :322)."""
    return F.layer_norm(x, (x.shape[-1],), _w(sd, name + ".weight", x.dtype), _w(sd, name + ".bias", x.dtype), 1e-12)


def gelu(x: torch.Tensor) -> torch.Tensor:
    """Exact erf GELU (models/vilbert_dialog.py:115-121)."""
    return x * 0.5 * (1.0 + torch.erf(x / math.sqrt(2.0)))


def additive_mask(mask: torch.Tensor, dtype) -> torch.Tensor:
    """(1 - m) * -10000 in fp32 as the reference does (models/vilbert_dialog.py:1415-1431)."""
    return ((1.0 - mask.to(torch.float32)) * MASK_NEG).to(dtype)


# --------------------------------------------------------------------------- embeddings
def _nodrop(site, x):
    return x


def text_embeddings(sd, cfg, input_ids, token_type_ids, position_ids, dtype, drop=_nodrop) -> torch.Tensor:
    """BertEmbeddingsDialog.forward (models/vilbert_dialog.py:326-356).

    Segment ids >= type_vocab_size select ``token_type_embeddings_extension[id - type_vocab_size]``,
    others ``token_type_embeddings[id]`` (:337-350).  ``sep_embeddings`` and the sinusoid table
    ``pe`` are never used in the forward.
    """
    p = "bert.embeddings."
    words = F.embedding(input_ids, _w(sd, p + "word_embeddings.weight", dtype))
    pos = F.embedding(position_ids, _w(sd, p + "position_embeddings.weight", dtype))
    is_ext = token_type_ids >= cfg.type_vocab_size
    base_ids = torch.where(is_ext, torch.zeros_like(token_type_ids), token_type_ids)
    ext_ids = torch.where(is_ext, token_type_ids - cfg.type_vocab_size, torch.zeros_like(token_type_ids))
    base = F.embedding(base_ids, _w(sd, p + "token_type_embeddings.weight", dtype))
    ext = F.embedding(ext_ids, _w(sd, p + "token_type_embeddings_extension.weight", dtype))
    types = torch.where(is_ext.unsqueeze(-1), ext, base)
    return drop("emb.txt", layer_norm(sd, p + "LayerNorm", words + pos + types))                  # :354-355


def image_embeddings(sd, image_feat, image_loc, drop=_nodrop) -> torch.Tensor:
    """BertImageEmbeddings.forward (models/vilbert_dialog.py:1487-1493)."""
    p = "bert.v_embeddings."
    return drop("emb.img", layer_norm(sd, p + "LayerNorm",
                                      linear(sd, p + "image_embeddings", image_feat) + linear(sd, p + "image_location_embeddings", image_loc)))


# --------------------------------------------------------------------------- attention
def _split_heads(x: torch.Tensor, heads: int) -> torch.Tensor:
    b, s, h = x.shape
    return x.view(b, s, heads, h // heads).permute(0, 2, 1, 3)


def _merge_heads(x: torch.Tensor) -> torch.Tensor:
    b, h, s, d = x.shape
    return x.permute(0, 2, 1, 3).reshape(b, s, h * d)


def attention(q, k, v, heads: int, add_mask: Optional[torch.Tensor], drop=_nodrop, site: str = "") -> torch.Tensor:
    """softmax(Q K^T / sqrt(d) + mask) V; scale before mask, dropout on the probabilities (models/vilbert_dialog.py:395-410)."""
    qh, kh, vh = _split_heads(q, heads), _split_heads(k, heads), _split_heads(v, heads)
    scores = torch.matmul(qh, kh.transpose(-1, -2)) / math.sqrt(qh.shape[-1])
    if add_mask is not None:
        scores = scores + add_mask
    probs = drop(site, torch.softmax(scores, dim=-1))
    return _merge_heads(torch.matmul(probs, vh))


def transformer_layer(sd, p: str, x, add_mask, heads: int, drop=_nodrop) -> torch.Tensor:
    """BertLayer / BertImageLayer (models/vilbert_dialog.py:479-483, :608-612)."""
    a = p + "attention."
    ctx = attention(linear(sd, a + "self.query", x), linear(sd, a + "self.key", x), linear(sd, a + "self.value", x),
                    heads, add_mask, drop, a + "probs")
    att = layer_norm(sd, a + "output.LayerNorm", drop(a + "output", linear(sd, a + "output.dense", ctx)) + x)       # :422-426
    inter = gelu(linear(sd, p + "intermediate.dense", att))                                      # :452-455
    return layer_norm(sd, p + "output.LayerNorm", drop(p + "output", linear(sd, p + "output.dense", inter)) + att)   # :465-469


def connection_layer(sd, cfg, p: str, img, img_add_mask, txt, co_add_mask, drop=_nodrop):
    """BertConnectionLayer.forward (models/vilbert_dialog.py:770-783).

    BertBiAttention (:655-723): stream 1 = image, stream 2 = text.  Text queries attend image keys
    with the image padding mask only (``attended_all_tensor1=True`` → co-mask not applied, :686-687);
    image queries attend text keys with the co-attention mask only (text padding mask is commented
    out, :705-709).  BertBiOutput (:744-754) swaps the contexts: image rows receive the
    image-queries-over-text context through dense1/LayerNorm1, text rows the other through dense2.
    """
    b = p + "biattention."
    heads = cfg.bi_num_attention_heads
    q1, k1, v1 = linear(sd, b + "query1", img), linear(sd, b + "key1", img), linear(sd, b + "value1", img)
    q2, k2, v2 = linear(sd, b + "query2", txt), linear(sd, b + "key2", txt), linear(sd, b + "value2", txt)
    ctx_txt_over_img = attention(q2, k1, v1, heads, img_add_mask, drop, b + "probs1")      # context_layer1 [B,S,Hb], dropout1 :693
    ctx_img_over_txt = attention(q1, k2, v2, heads, co_add_mask, drop, b + "probs2")       # context_layer2 [B,R,Hb], dropout2 :716
    o = p + "biOutput."
    img_att = layer_norm(sd, o + "LayerNorm1", drop(o + "dense1", linear(sd, o + "dense1", ctx_img_over_txt)) + img)    # :745-747
    txt_att = layer_norm(sd, o + "LayerNorm2", drop(o + "dense2", linear(sd, o + "dense2", ctx_txt_over_img)) + txt)    # :748-750
    img_out = layer_norm(sd, p + "v_output.LayerNorm",
                         drop(p + "v_output", linear(sd, p + "v_output.dense", gelu(linear(sd, p + "v_intermediate.dense", img_att)))) + img_att)
    txt_out = layer_norm(sd, p + "t_output.LayerNorm",
                         drop(p + "t_output", linear(sd, p + "t_output.dense", gelu(linear(sd, p + "t_intermediate.dense", txt_att)))) + txt_att)
    return img_out, txt_out


def layer_schedule(cfg):
    """Execution order of BertEncoder.forward (models/vilbert_dialog.py:842-929)."""
    order, vs, ts = [], 0, 0
    for c, (ve, te) in enumerate(zip(cfg.v_biattention_id, cfg.t_biattention_id)):
        order += [("v", i) for i in range(vs, ve)] + [("t", i) for i in range(ts, te)] + [("c", c)]
        vs, ts = ve, te
    order += [("v", i) for i in range(vs, cfg.v_num_hidden_layers)]
    order += [("t", i) for i in range(ts, cfg.num_hidden_layers)]
    return order


def encoder(sd, cfg, txt, img, txt_add_mask, img_add_mask, co_add_mask, taps: Optional[dict] = None, drop=_nodrop):
    for kind, i in layer_schedule(cfg):
        if kind == "t":
            txt = transformer_layer(sd, f"bert.encoder.layer.{i}.", txt, txt_add_mask, cfg.num_attention_heads, drop)
        elif kind == "v":
            img = transformer_layer(sd, f"bert.encoder.v_layer.{i}.", img, img_add_mask, cfg.v_num_attention_heads, drop)
        else:
            img, txt = connection_layer(sd, cfg, f"bert.encoder.c_layer.{i}.", img, img_add_mask, txt, co_add_mask, drop)
        if taps is not None:
            taps[f"{kind}{i}.txt"] = txt
            taps[f"{kind}{i}.img"] = img
    return txt, img


# --------------------------------------------------------------------------- heads
def lm_transform(sd, rows: torch.Tensor) -> torch.Tensor:
    """BertPredictionHeadTransform (models/vilbert_dialog.py:982-986)."""
    t = "cls.predictions.transform."
    return layer_norm(sd, t + "LayerNorm", gelu(linear(sd, t + "dense", rows)))


def lm_logits(sd, rows: torch.Tensor) -> torch.Tensor:
    """BertLMPredictionHead.forward (models/vilbert_dialog.py:1023-1026); decoder tied to word embeddings."""
    h = lm_transform(sd, rows)
    return F.linear(h, _w(sd, "cls.predictions.decoder.weight", h.dtype)) + _w(sd, "cls.predictions.bias", h.dtype)


def nsp_logits(sd, txt, img, drop=_nodrop) -> torch.Tensor:
    """Poolers (:946-967) and the 'mul' fusion NSP head with its dropout on the fused vector (:1062-1070)."""
    pt = torch.relu(linear(sd, "bert.t_pooler.dense", txt[:, 0]))
    pv = torch.relu(linear(sd, "bert.v_pooler.dense", img[:, 0]))
    return linear(sd, "cls.bi_seq_relationship", drop("nsp.pooled", pt * pv))


def image_logits(sd, img) -> torch.Tensor:
    """BertImagePredictionHead (models/vilbert_dialog.py:1085-1088)."""
    t = "cls.imagePredictions."
    h = layer_norm(sd, t + "transform.LayerNorm", gelu(linear(sd, t + "transform.dense", img)))
    return linear(sd, t + "decoder", h)


# --------------------------------------------------------------------------- full forward
def forward(sd: Dict[str, torch.Tensor], cfg, input_ids, image_feat, image_loc, token_type_ids, position_ids,
            attention_mask, image_attention_mask, co_attention_mask,
            masked_lm_labels=None, next_sentence_label=None, image_label=None, image_target=None,
            nsp_weight=None, lm_weight=None, dtype=torch.float32, full_logits: bool = False,
            taps: Optional[dict] = None, drop=_nodrop) -> Dict[str, torch.Tensor]:
    """BertModel.forward + BertForMultiModalPreTraining.forward (models/vilbert_dialog.py:1359-1472, :1519-1626).

    Returns a dict: ``sequence_output_t`` [B,S,H], ``sequence_output_v``, ``nsp_scores`` [B,2],
    ``token_rows`` [n,2] (b,s) of positions with label != -1, ``token_logp`` [n] = log_softmax at the
    label, ``token_ul`` [n] = log(clamp(1-p,1e-6)) at the label, ``seq_score`` [B] = val_lm.py:131-136's
    ``-nll.sum(-1)``, ``prediction_scores_t`` when ``full_logits`` (the reference's full-vocab output),
    and, when the three training targets are present (:1559), ``lm_loss`` / ``img_loss`` / ``nsp_loss``.
    """
    image_feat, image_loc = image_feat.to(dtype), image_loc.to(dtype)
    if attention_mask.dim() == 3:                                   # :1396-1399
        txt_add = additive_mask(attention_mask[:, None, :, :], dtype)
    else:
        txt_add = additive_mask(attention_mask[:, None, None, :], dtype)
    img_add = additive_mask(image_attention_mask[:, None, None, :], dtype)     # :1405-1406,1423
    co_add = additive_mask(co_attention_mask[:, None, :, :], dtype)            # :1427-1431

    txt = text_embeddings(sd, cfg, input_ids, token_type_ids, position_ids, dtype, drop)
    img = image_embeddings(sd, image_feat, image_loc, drop)
    if taps is not None:
        taps["emb.txt"], taps["emb.img"] = txt, img
    txt, img = encoder(sd, cfg, txt, img, txt_add, img_add, co_add, taps, drop)
    out = {"sequence_output_t": txt, "sequence_output_v": img, "nsp_scores": nsp_logits(sd, txt, img, drop)}

    B, S, _ = txt.shape
    if full_logits:
        logits = lm_logits(sd, txt)                                 # [B,S,V] as the reference materialises
        out["prediction_scores_t"] = logits
    if masked_lm_labels is not None:
        rows = (masked_lm_labels != -1).nonzero()                   # [n,2]
        labels = masked_lm_labels[rows[:, 0], rows[:, 1]]
        if full_logits:
            row_logits = logits[rows[:, 0], rows[:, 1]]
        else:
            row_logits = lm_logits(sd, txt[rows[:, 0], rows[:, 1]])
        logp_all = torch.log_softmax(row_logits, dim=-1)
        logp = logp_all.gather(1, labels[:, None])[:, 0]
        p_all = torch.softmax(row_logits, dim=-1)
        ul = torch.log(torch.clamp(1.0 - p_all, min=1e-6)).gather(1, labels[:, None])[:, 0]   # :1587
        seq = torch.zeros(B, dtype=logp.dtype, device=logp.device).index_add_(0, rows[:, 0], logp)      # val_lm.py:131-136
        out.update(token_rows=rows, token_logp=logp, token_ul=ul, seq_score=seq)

    if masked_lm_labels is not None and next_sentence_label is not None and image_target is not None:
        # image KL loss (:1569-1574)
        v_logits = image_logits(sd, img)
        kl = F.kl_div(torch.log_softmax(v_logits, dim=2), image_target.to(dtype), reduction="none")
        sel = (image_label == 1)
        out["img_loss"] = (kl * sel.unsqueeze(2).to(dtype)).sum() / max(sel.sum(), 0)
        # likelihood / unlikelihood (:1577-1595)
        if lm_weight is not None:
            w_all = lm_weight.reshape(-1)
            lab_all = masked_lm_labels.reshape(-1)
            # the reference selects rows by weight, then nll_loss(ignore_index=-1) drops label -1 rows
            l_sel = (w_all > 0) & (lab_all != -1)
            ul_sel = (w_all == -1) & (lab_all != -1)
            flat = rows[:, 0] * S + rows[:, 1]
            pos_of = torch.full((B * S,), -1, dtype=torch.long)
            pos_of[flat] = torch.arange(flat.numel())
            l_idx, ul_idx = pos_of[l_sel.nonzero()[:, 0]], pos_of[ul_sel.nonzero()[:, 0]]
            l_loss = (-logp[l_idx] * w_all[l_sel].to(dtype)).sum()
            ul_loss = (-ul[ul_idx]).sum()
            out["lm_loss"] = (l_loss + ul_loss) / (lm_weight != 0).sum()
        else:
            out["lm_loss"] = (-logp).mean()                         # CrossEntropyLoss(ignore_index=-1), :1601-1604
        # NSP weighted CE (:1605-1621)
        if nsp_weight is None:
            nw = torch.ones(2, dtype=dtype)
        else:
            nw = nsp_weight.reshape(-1)[:2].to(dtype)
        nw = nw / nw[0]
        out["nsp_loss"] = F.cross_entropy(out["nsp_scores"].view(-1, 2), next_sentence_label.view(-1), weight=nw,
                                          reduction="mean")
    return out


def score_candidates(sd, cfg, batch: dict, chunk: int = 25, dtype=torch.float32, full_logits: bool = True):
    """val_lm.py:104-139 restated: chunked forward, per-sequence sum of answer-token log-probs.

    ``batch`` holds the flattened per-sequence tensors (tokens, segments, positions, mask (labels),
    txt_attention_mask, co_attention_mask, image_feat, image_loc, image_mask).  Returns
    (seq_score [N], nsp_scores [N,2], token_logp list).
    """
    n = batch["tokens"].shape[0]
    scores, nsp, tok = [], [], []
    with torch.no_grad():
        for s in range(0, n, chunk):
            e = min(n, s + chunk)
            o = forward(sd, cfg, batch["tokens"][s:e], batch["image_feat"][s:e], batch["image_loc"][s:e],
                        batch["segments"][s:e], batch["positions"][s:e], batch["txt_attention_mask"][s:e],
                        batch["image_mask"][s:e], batch["co_attention_mask"][s:e],
                        masked_lm_labels=batch["mask"][s:e], dtype=dtype, full_logits=full_logits)
            scores.append(o["seq_score"])
            nsp.append(o["nsp_scores"])
            tok.append(o["token_logp"])
    return torch.cat(scores), torch.cat(nsp), torch.cat(tok)
