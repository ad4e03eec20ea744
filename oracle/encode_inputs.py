"""ORACLE (test infrastructure): restatement of the reference's sequence/mask encoders.

Follows ``/root/reference/utils/data_utils.py``: ``encode_input_gen`` (:139-288), ``encode_input_dis``
(:291-428), ``encode_input`` (:430-436), ``encode_image_input`` (:438-482, deterministic part) and the
way ``dataloader/dataloader_visdial.py:322-457`` assembles a validation item.  Written from the
behaviour (the dense masks are produced from closed-form row intervals rather than by the
reference's slice assignments); ``tests/golden/make_golden.py`` asserts, in the build container,
that every tensor equals the reference functions' output on the same numpy RNG stream, and the
committed fixtures pin it on the GPU box.

Random draws follow the reference's order on the numpy global-style stream so a ``RandomState(seed)``
here reproduces ``np.random.seed(seed)`` there.
"""
from __future__ import annotations

from typing import List, Sequence

import numpy as np
import torch

CLS, SEP, MASK = 101, 102, 103          # bert-base-uncased ids (dataloader_visdial.py:58-62)


def _pad(values: Sequence, n: int) -> torch.Tensor:
    """list2tensorpad (data_utils.py:58-63): LongTensor (floats truncate toward zero), zero padded."""
    t = torch.LongTensor([list(values)])
    out = torch.zeros(1, n, dtype=torch.long)
    out[0, : t.shape[1]] = t
    return out


def gen_mask_rows(ctx: int, L: int, last_len: int, S: int) -> torch.Tensor:
    """Dense [S,S] bool mask of the generative mode from its closed form (data_utils.py:149-210).

    row 0 → [0,T); rows [1,ctx) → [1,ctx); A rows i∈[ctx,L) → [1,i]; B rows j∈[L,T) → [1,j-last_len) ∪ {j};
    rows ≥ T → empty.  T = L + last_len; everything clipped to S.
    """
    T = L + last_len
    r = torch.arange(S)[:, None]
    c = torch.arange(S)[None, :]
    m = torch.zeros(S, S, dtype=torch.bool)
    m |= (r == 0) & (c < T)
    m |= (r >= 1) & (r < ctx) & (c >= 1) & (c < ctx)
    m |= (r >= ctx) & (r < L) & (c >= 1) & (c <= r)
    m |= (r >= L) & (r < T) & (((c >= 1) & (c < r - last_len)) | (c == r))
    return m


def encode_gen(utterances: List[List[int]], start_segment: int, cls=CLS, sep=SEP, mask_id=MASK, max_seq_len=256,
               max_sep_len=25, mask_prob=0.1, is_negative=0, weight=1, vocab_size=None, rng=np.random, _dis=False):
    seg = start_segment
    toks, segs, poss, seps, msk, wts = [cls], [seg], [0], [], [0], [0]
    n_utt = len(utterances)
    L = last_len = 0
    for k, utt in enumerate(utterances, start=1):
        n = len(utt)
        last = (k == n_utt)
        draws = [0] * n if (last and n <= 1) else [1 if rng.rand() < mask_prob else 0 for _ in range(n)]
        toks += list(utt) + [sep]
        segs += [seg] * (n + 1)
        msk += draws + [0]
        wts += ([0] * n if (last and is_negative) else draws) + [0]
        span = list(range(len(poss), len(poss) + n + 1))
        poss += span
        seps.append((seps[-1] if seps else 0) + n + 1)
        if last:
            last_len, L = n + 1, len(toks)
            if not _dis:                                    # the masked copy of the answer (B rows)
                toks += list(utt) + [sep]
                segs += [seg] * (n + 1)
                msk += [1] * (n + 1)
                wts += [(-weight if is_negative else weight)] * (n + 1)
                poss += span                                # same position ids as the visible copy (:227)
                seps.append(seps[-1] + n + 1)
        seg ^= 1
    assert len(segs) == len(toks) == len(msk) == seps[-1] + 1
    if len(toks) > max_seq_len:
        toks, segs, poss, msk, wts = (x[:max_seq_len] for x in (toks, segs, poss, msk, wts))
        seps[-1] = max_seq_len - 1
    tokens = _pad(toks, max_seq_len)
    labels = _pad(msk, max_seq_len)
    sel = labels[0] == 1
    labels[0, ~sel] = -1
    labels[0, sel] = tokens[0, sel]
    tokens[0, sel] = mask_id
    for pos in torch.nonzero(sel)[:, 0].tolist():           # 80/10/10 rule as the reference evaluates it (:252-257)
        if rng.rand() < 0.8 or vocab_size is None or pos >= L:
            continue
        if rng.rand() < 0.5:
            tokens[0, pos] = rng.randint(0, vocab_size)
    S = max_seq_len
    if _dis:
        r = torch.arange(S)
        att = ((r[:, None] < L) & (r[None, :] < L)).long()
        co = (r < L).long()
    else:
        att = gen_mask_rows(L - last_len, L, last_len, S)
        r = torch.arange(S)
        co = ((r >= 1) & (r < L - last_len)).long()
    return (tokens, _pad(segs, S), _pad(poss, S), _pad(seps, max_sep_len), labels, _pad(wts, S),
            att.unsqueeze(0), co.unsqueeze(0))


def encode_dis(utterances, start_segment, cls=CLS, sep=SEP, mask_id=MASK, **kw):
    return encode_gen(utterances, start_segment, cls, sep, mask_id, _dis=True, **kw)


def encode(dis_rate, utterances, start_segment, rng=np.random, **kw):
    """encode_input (data_utils.py:430-436): one uniform draw picks the mode."""
    if rng.rand() < dis_rate:
        return encode_dis(utterances, start_segment, rng=rng, **kw)
    return encode_gen(utterances, start_segment, rng=rng, **kw)


# --------------------------------------------------------------------------- synthetic dialogs
def synth_round(rng: np.random.RandomState, n_candidates=100, caption_len=20, n_hist_utts=19, utt_len=10,
                question_len=7, ans_len_range=(1, 7), vocab=(1000, 30522)):
    """Config-1 dialog round: shared context + ``n_candidates`` answer options (token-id lists)."""
    draw = lambda n: rng.randint(vocab[0], vocab[1], size=n).tolist()
    context = [draw(caption_len)] + [draw(utt_len) for _ in range(n_hist_utts)] + [draw(question_len)]
    answers = [draw(int(rng.randint(ans_len_range[0], ans_len_range[1] + 1))) for _ in range(n_candidates)]
    return context, answers


def synth_image(rng: np.random.RandomState, n_boxes=36, feat_dim=2048):
    """36 region features + prepended global mean / whole-image box (image_features_reader.py:85-88,102)."""
    feats = rng.randn(n_boxes, feat_dim).astype(np.float32)
    feats = np.concatenate([feats.mean(0, keepdims=True), feats], 0)
    loc = rng.rand(n_boxes + 1, 5).astype(np.float32)
    loc[0] = [0, 0, 1, 1, 1]
    return torch.from_numpy(feats), torch.from_numpy(loc), torch.ones(n_boxes + 1)


def build_batch(context, answers, feats, loc, image_mask, mode="gen", start_segment=1, encoder=None, **kw):
    """Flattened per-sequence batch the way val_lm.py:55-93 hands it to train.forward."""
    enc = encoder or (encode_gen if mode == "gen" else encode_dis)
    cols = [[] for _ in range(8)]
    for ans in answers:
        out = enc(context + [ans], start_segment, mask_prob=0, **kw)
        for c, o in zip(cols, out):
            c.append(o)
    tokens, segments, positions, sep_indices, labels, weights, att, co = (torch.cat(c, 0) for c in cols)
    n, R = tokens.shape[0], feats.shape[0]
    return {
        "tokens": tokens, "segments": segments, "positions": positions, "sep_indices": sep_indices,
        "mask": labels, "weights": weights, "txt_attention_mask": att,
        "co_attention_mask": co.unsqueeze(1).repeat(1, R, 1),           # dataloader_visdial.py:455
        "image_feat": feats.unsqueeze(0).expand(n, -1, -1).contiguous(),
        "image_loc": loc.unsqueeze(0).expand(n, -1, -1).contiguous(),
        "image_mask": image_mask.unsqueeze(0).expand(n, -1).contiguous(),
    }
