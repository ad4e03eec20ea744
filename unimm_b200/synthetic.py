"""Synthetic VisDial-shaped inputs for the bench and the property tests (no datasets offline).

Builds exactly what the reference's val pipeline would hand to the encoder for generative ranking
(reference dataloader/dataloader_visdial.py:322-457 driving utils/data_utils.py:encode_input_gen with
``mask_prob=0``), but directly as token/segment/position/label arrays plus the 4-integer descriptor —
no dense ``[S,S]`` mask is ever materialised.  Layout of one sequence (S = 256):

    [CLS] caption [SEP] u1 [SEP] ... question [SEP] | answer [SEP] | [MASK]*len(answer) [MASK] | pad
    └──────────── context rows [0,ctx) ───────────┘ └ A: [ctx,L) ┘ └────── B: [L,T) ──────┘

labels are -1 except on the B copy (answer tokens + [SEP]); B repeats A's position ids; segments
alternate per utterance starting at 1 (reference pruneRounds / encode_input_gen).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Sequence

import numpy as np
import torch

CLS, SEP, MASK = 101, 102, 103
S_MAX = 256


@dataclass
class Round:
    """One (image, round) unit: n candidate sequences that share context and image."""
    tokens: np.ndarray      # [n,S] int64
    segments: np.ndarray    # [n,S] int64
    positions: np.ndarray   # [n,S] int64
    labels: np.ndarray      # [n,S] int64, -1 = ignore
    desc: np.ndarray        # [n,4] int32 (mode, ctx, L, last_len)


def encode_round_gen(context: Sequence[Sequence[int]], answers: Sequence[Sequence[int]], start_segment: int = 1,
                     S: int = S_MAX) -> Round:
    """All candidates of a round at once (index arithmetic over [n, S]); ``_encode_round_gen_loop`` is the per-candidate
    statement of the same layout and tests/test_host_cpu.py requires the two to agree."""
    n = len(answers)
    ctx_tok, ctx_seg = [CLS], [start_segment]
    seg = start_segment
    for u in context:
        ctx_tok += list(u) + [SEP]
        ctx_seg += [seg] * (len(u) + 1)
        seg ^= 1
    ctx = len(ctx_tok)
    last = np.asarray([len(a) + 1 for a in answers], np.int64)
    if n and ctx + 2 * int(last.max()) > S:
        raise ValueError("synthetic sequence exceeds max_seq_len; shorten the dialog")
    mx = int(last.max()) if n else 1
    ans = np.full((n, mx), SEP, np.int64)                 # answer tokens followed by [SEP]
    for j, a in enumerate(answers):
        ans[j, :len(a)] = a
    col = np.arange(S)[None, :]
    L, T = (ctx + last)[:, None], (ctx + 2 * last)[:, None]
    in_a, in_b = (col >= ctx) & (col < L), (col >= L) & (col < T)
    k_a = np.clip(col - ctx, 0, mx - 1)
    k_b = np.clip(col - L, 0, mx - 1)
    rows = np.arange(n)[:, None]
    tokens = np.zeros((n, S), np.int64)
    tokens[:, :ctx] = ctx_tok
    tokens = np.where(in_a, ans[rows, k_a], np.where(in_b, MASK, tokens))
    segments = np.zeros((n, S), np.int64)
    segments[:, :ctx] = ctx_seg
    segments = np.where(in_a | in_b, seg, segments)
    positions = np.where(col < L, col, np.where(in_b, col - last[:, None], 0)) + np.zeros((n, 1), np.int64)
    labels = np.where(in_b, ans[rows, k_b], -1)
    desc = np.stack([np.zeros(n, np.int64), np.full(n, ctx), ctx + last, last], 1).astype(np.int32)
    return Round(np.ascontiguousarray(tokens), np.ascontiguousarray(segments), np.ascontiguousarray(positions.astype(np.int64)),
                 np.ascontiguousarray(labels.astype(np.int64)), desc)


def _encode_round_gen_loop(context: Sequence[Sequence[int]], answers: Sequence[Sequence[int]], start_segment: int = 1,
                           S: int = S_MAX) -> Round:
    n = len(answers)
    tokens = np.zeros((n, S), np.int64)
    segments = np.zeros((n, S), np.int64)
    positions = np.zeros((n, S), np.int64)
    labels = np.full((n, S), -1, np.int64)
    desc = np.zeros((n, 4), np.int32)
    ctx_tok, ctx_seg = [CLS], [start_segment]
    seg = start_segment
    for u in context:
        ctx_tok += list(u) + [SEP]
        ctx_seg += [seg] * (len(u) + 1)
        seg ^= 1
    ctx = len(ctx_tok)
    ctx_tok, ctx_seg = np.asarray(ctx_tok, np.int64), np.asarray(ctx_seg, np.int64)
    for j, ans in enumerate(answers):
        a = np.asarray(list(ans) + [SEP], np.int64)
        last = len(a)
        L, T = ctx + last, ctx + 2 * last
        if T > S:
            raise ValueError("synthetic sequence exceeds max_seq_len; shorten the dialog")
        tokens[j, :ctx] = ctx_tok
        tokens[j, ctx:L] = a
        tokens[j, L:T] = MASK
        segments[j, :ctx] = ctx_seg
        segments[j, ctx:T] = seg
        positions[j, :L] = np.arange(L)
        positions[j, L:T] = np.arange(ctx, L)
        labels[j, L:T] = a
        desc[j] = (0, ctx, L, last)
    return Round(tokens, segments, positions, labels, desc)


def encode_round_dis(context, answers, start_segment: int = 1, S: int = S_MAX) -> Round:
    """Discriminative layout (encode_input_dis with mask_prob=0): no masked copy, no labels."""
    n = len(answers)
    r = Round(np.zeros((n, S), np.int64), np.zeros((n, S), np.int64), np.zeros((n, S), np.int64),
              np.full((n, S), -1, np.int64), np.zeros((n, 4), np.int32))
    ctx_tok, ctx_seg = [CLS], [start_segment]
    seg = start_segment
    for u in context:
        ctx_tok += list(u) + [SEP]
        ctx_seg += [seg] * (len(u) + 1)
        seg ^= 1
    ctx = len(ctx_tok)
    for j, ans in enumerate(answers):
        a = list(ans) + [SEP]
        L = ctx + len(a)
        r.tokens[j, :ctx], r.tokens[j, ctx:L] = ctx_tok, a
        r.segments[j, :ctx], r.segments[j, ctx:L] = ctx_seg, seg
        r.positions[j, :L] = np.arange(L)
        r.desc[j] = (1, 0, L, 0)
    return r


def _draw(rng, n, vocab=(1000, 30522)) -> List[int]:
    return rng.randint(vocab[0], vocab[1], size=n).tolist()


def synth_context(rng: np.random.RandomState, round_id: int = 10, caption_len: int = 20, utt_len: int = 10,
                  question_len: int = 7) -> List[List[int]]:
    """Caption + (round_id-1) question/answer pairs + the current question.  ``round_id=10`` with one extra
    history utterance is SURVEY.md §8(d)'s config 1 (239 context positions)."""
    hist = 2 * (round_id - 1) + (1 if round_id == 10 else 0)
    return [_draw(rng, caption_len)] + [_draw(rng, utt_len) for _ in range(hist)] + [_draw(rng, question_len)]


def synth_answers(rng: np.random.RandomState, n: int = 100, len_range=(1, 7)) -> List[List[int]]:
    return [_draw(rng, int(rng.randint(len_range[0], len_range[1] + 1))) for _ in range(n)]


def synth_image(rng: np.random.RandomState, n_boxes: int = 36, feat_dim: int = 2048):
    """[37,2048] features with the global mean row prepended, [37,5] boxes, [37] mask
    (reference utils/image_features_reader.py:85-88,102)."""
    f = rng.randn(n_boxes, feat_dim).astype(np.float32)
    f = np.concatenate([f.mean(0, keepdims=True), f], 0)
    loc = rng.rand(n_boxes + 1, 5).astype(np.float32)
    loc[0] = [0, 0, 1, 1, 1]
    return f, loc, np.ones(n_boxes + 1, np.float32)


def synth_dialog_rounds(image_id: int, rounds: Sequence[int] = tuple(range(1, 11)), n_candidates: int = 100, mode: str = "gen"):
    """All requested rounds of one synthetic image (seed = image id): (image arrays, [Round, ...]).  ``mode="dis"``: the same
    dialogs under the discriminative layout (val.py's NSP ranking)."""
    rng = np.random.RandomState(100003 + image_id)
    img = synth_image(rng)
    out = []
    enc = encode_round_gen if mode == "gen" else encode_round_dis
    for r in rounds:
        out.append(enc(synth_context(rng, r), synth_answers(rng, n_candidates)))
    # like the reference's loader (dataloader_visdial.py:437-457: one [rounds * options, 256] tensor per field and image), keep
    # the rounds of an image as row ranges of ONE array per field: the packer then gathers per image instead of per round
    cat = {f: np.concatenate([getattr(r, f) for r in out], 0) for f in ("tokens", "segments", "positions", "labels", "desc")}
    views, s = [], 0
    for r in out:
        e = s + len(r.tokens)
        views.append(Round(cat["tokens"][s:e], cat["segments"][s:e], cat["positions"][s:e], cat["labels"][s:e], cat["desc"][s:e]))
        s = e
    return img, views


def stack_rounds(rounds: Sequence[Round]):
    """Concatenate units into flat torch tensors + a feat_index (sequence -> unit) vector."""
    cat = lambda f: torch.from_numpy(np.concatenate([getattr(r, f) for r in rounds], 0))
    index = np.concatenate([np.full(len(r.tokens), u, np.int32) for u, r in enumerate(rounds)])
    return cat("tokens"), cat("segments"), cat("positions"), cat("labels"), cat("desc"), torch.from_numpy(index)


# ------------------------------------------------------------------------------------------------ configs 3 / 4 / 5 (bench inputs)
def _context_arrays(context, start_segment=1):
    tok, seg, s = [CLS], [start_segment], start_segment
    for u in context:
        tok += list(u) + [SEP]
        seg += [s] * (len(u) + 1)
        s ^= 1
    return tok, seg, s


def encode_train_sequence(rng, context, answer, dis: bool, negative: bool, mask_prob: float = 0.15, weight: int = 1, S: int = S_MAX,
                          vocab: int = 30522):
    """One training sequence the way ``encode_input`` shapes it (utils/data_utils.py:139-436): random 15 % masking with the
    80/10/10 rule on the visible tokens, the masked answer copy in generative mode, token weights +w (likelihood), -w on a
    negative's masked copy (unlikelihood) and 0 on a negative's visible answer.  Shape-faithful synthetic data for the bench (the
    parity fixtures use the reference's own encoder)."""
    tok, seg, cur = _context_arrays(context)
    ctx = len(tok)
    a = list(answer) + [SEP]
    last = len(a)
    tokens = np.zeros(S, np.int64); segments = np.zeros(S, np.int64); positions = np.zeros(S, np.int64)
    labels = np.full(S, -1, np.int64); weights = np.zeros(S, np.int64)
    L = ctx + last
    tokens[:ctx], tokens[ctx:L] = tok, a
    segments[:ctx], segments[ctx:L] = seg, cur
    positions[:L] = np.arange(L)
    # random masking of the visible (non-special) tokens
    special = np.zeros(S, bool)
    special[0] = True
    special[:L] |= tokens[:L] == SEP
    draw = (rng.rand(L) < mask_prob) & ~special[:L]
    if last <= 2:
        draw[ctx:L] = False                                   # one-token answers are never masked in place (:173-175)
    for p in np.flatnonzero(draw):
        labels[p] = tokens[p]
        weights[p] = 0 if (negative and p >= ctx) else 1
        r = rng.rand()
        tokens[p] = MASK if r < 0.8 else (rng.randint(0, vocab) if r < 0.9 else tokens[p])
    if dis:
        desc = (1, 0, L, 0)
    else:
        T = L + last
        if T > S:
            raise ValueError("synthetic training sequence exceeds max_seq_len")
        tokens[L:T] = MASK
        segments[L:T] = cur
        positions[L:T] = np.arange(ctx, L)
        labels[L:T] = a
        weights[L:T] = -weight if negative else weight
        desc = (0, ctx, L, last)
    return tokens, segments, positions, labels, weights, np.asarray(desc, np.int32)


def train_batch(seed: int, n_images: int = 40, per_image: int = 6, dis_rate: float = 0.5, mask_prob: float = 0.15):
    """BASELINE config 3's batch: ``n_images`` x (1 positive + ``per_image``-1 negatives of one round), mode drawn per sequence.
    Returns a dict of numpy arrays: ids [B,256] x3, labels, weights, desc [B,4], next_sentence_label [B], seq_image [B] and the
    per-image blocks image_feat / image_loc / image_mask / image_label [n,37] / image_target [n,37,1601]."""
    rng = np.random.RandomState(seed)
    cols = [[] for _ in range(6)]
    nsl, seq_image, feats, locs, masks, ilabels, targets = [], [], [], [], [], [], []
    for i in range(n_images):
        f, l, m = synth_image(rng)
        il = np.where(rng.rand(37) < 0.15, 1, -1)
        il[0] = 0
        f = f.copy()
        f[(il == 1) & (rng.rand(37) < 0.9)] = 0
        t = rng.rand(37, 1601).astype(np.float32) ** 8
        t /= t.sum(-1, keepdims=True)
        feats.append(f), locs.append(l), masks.append(m), ilabels.append(il), targets.append(t)
        context = synth_context(rng, int(rng.randint(1, 11)))
        for j in range(per_image):
            out = encode_train_sequence(rng, context, _draw(rng, int(rng.randint(1, 8))), dis=rng.rand() < dis_rate, negative=j > 0,
                                        mask_prob=mask_prob)
            for c, o in zip(cols, out):
                c.append(o)
            nsl.append(int(j > 0)), seq_image.append(i)
    tokens, segments, positions, labels, weights, desc = (np.stack(c) for c in cols)
    return {"tokens": tokens, "segments": segments, "positions": positions, "labels": labels, "weights": weights, "desc": desc,
            "next_sentence_label": np.asarray(nsl, np.int64), "seq_image": np.asarray(seq_image, np.int32),
            "image_feat": np.stack(feats), "image_loc": np.stack(locs), "image_mask": np.stack(masks),
            "image_label": np.stack(ilabels).astype(np.int64), "image_target": np.stack(targets)}
