"""Prefix-shared ("packed") input layout for generative ranking.

In generative mode (reference utils/data_utils.py:199-210) the context rows ``[1, ctx)`` attend only the
context, and the image rows attend only the context (co-attention mask ``[1, ctx)``), so for the 100 candidate
answers of one dialog round they are bit-identical.  A *unit* = (image, round) therefore stores

    shared rows      original positions 1 .. ctx-1            (once per unit)
    candidate rows   [CLS] (position 0), the visible answer copy A (positions ctx .. L-1) and the masked copy B
                     (positions L .. T-1)                      (once per candidate: 1 + 2*last_len rows)

Text rows of a forward are packed as ``[all units' shared rows | all candidates' rows]``.  This module is the readable
SPECIFICATION of that layout (vectorised numpy); the sweep and the bench pack through ``unimm_b200.flat_packer`` (threaded C++ in
``csrc/packer.cu``, same arrays row for row — ``tests/test_packer_cpu.py`` — plus truncated sequences and one feature block per
image).  It turns the reference-format per-sequence arrays (tokens / segments / positions / labels ``[n,256]`` + descriptors, i.e. what
``encode_input_gen`` produces and ``unimm_b200.synthetic`` generates) into that layout plus the attention job
lists and per-row intervals ``libunimm_b200`` consumes (include/unimm_b200.h: ``unimm_packed_batch_t``).

Which keys a packed row attends (equals the dense mask restricted to real rows):
    shared row          all shared rows of its unit
    [CLS]               shared rows + all rows of its own candidate            (dense: [0, T))
    A_k                 shared rows + A_0..A_k                                  (dense: [1, i])
    B_k                 shared rows + A_0..A_{k-1} + itself                     (dense: [1, j-last) U {j})
    image region        the unit's image rows (self) / the unit's shared rows (co-attention)
    any text row        the unit's image rows in the text->image co-attention

``scores_only=True`` additionally drops the candidate rows that no labelled position can see: [CLS] (position 0 is outside
every other row's interval) and A_{last-1} (the visible copy's closing [SEP]: B_k sees A_0..A_{k-1} with k <= last-1, A_k sees
A_0..A_k).  They only feed the pooled NSP logit, which val_lm.py:124 fetches but never uses for the ranking, so a candidate
then costs 2*last_len - 1 rows instead of 2*last_len + 1 and the sequence log-likelihoods are unchanged
(tests/test_packing_cpu.py proves the closure on the dense masks; tests/test_parity_gpu.py compares the scores).
In that layout B_0 — the first masked position: [MASK] at position ctx, attending the context and itself — is a leaf that no
other kept row attends, and it is the SAME row for every candidate of the unit (``share_first_mask``, checked on the token /
segment / position arrays): it is stored once at the head of the unit's candidate block and listed in ``lm_rows`` once per
candidate (with that candidate's label), so a candidate costs 2*last_len - 2 own rows.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import List, Optional, Sequence

import numpy as np
import torch

R_DEFAULT = 37
Q_TILE = 128          # query rows per CTA of the candidate attention (csrc/attention_jobs.cu, NW = 8)


@dataclass
class UnitArrays:
    """One (image, round): dense reference-format arrays of its n candidates (all share the context)."""
    tokens: np.ndarray      # [n,S]
    segments: np.ndarray    # [n,S]
    positions: np.ndarray   # [n,S]
    labels: np.ndarray      # [n,S]
    desc: np.ndarray        # [n,4] (mode, ctx, L, last_len)
    image_slot: int = 0     # which feature block of the batch this unit uses


class PackedBatch:
    """Host-side packed batch (torch CPU tensors, optionally pinned) + counts."""

    INT_FIELDS = ("input_ids", "token_type_ids", "position_ids", "row_iv", "jobs_text_self", "jobs_t2i", "jobs_i2t",
                  "jobs_img_self", "lm_rows", "lm_labels", "cand_lm_off", "cand_cls_row", "cand_img_row", "lm_urows", "lm_uidx")
    FLOAT_FIELDS = ("image_feat", "image_loc", "image_mask")

    def __init__(self, **kw):
        self.__dict__.update(kw)

    def tensors(self):
        d = {k: getattr(self, k) for k in self.INT_FIELDS + self.FLOAT_FIELDS}
        if getattr(self, "unit_image", None) is not None:       # one feature block per IMAGE (flat_packer): unit -> block
            d["unit_image"] = self.unit_image
        return d

    def pin(self) -> "PackedBatch":
        for k, t in self.tensors().items():
            setattr(self, k, t.contiguous().pin_memory())
        return self

    def to(self, device) -> "PackedBatch":
        out = PackedBatch(**self.__dict__)
        for k, t in self.tensors().items():
            setattr(out, k, t.to(device, non_blocking=True))
        return out

    def bytes(self) -> int:
        return sum(t.numel() * t.element_size() for t in self.tensors().values())

    def c_struct(self):
        from ._lib import PackedBatchStruct
        s = PackedBatchStruct()
        s.n_units, s.n_cands, s.n_text_rows = self.n_units, self.n_cands, self.n_text_rows
        for k, t in self.tensors().items():
            setattr(s, "d_" + k, C.c_void_p(t.data_ptr()))
        s.n_jobs_text_self, s.max_q_text_self = self.jobs_text_self.shape[0], self.max_q_text_self
        s.n_jobs_text_ctx, s.cand_halo = self.n_jobs_text_ctx, self.cand_halo
        s.n_jobs_t2i, s.max_q_t2i = self.jobs_t2i.shape[0], self.max_q_t2i
        s.n_jobs_i2t, s.n_jobs_img_self = self.jobs_i2t.shape[0], self.jobs_img_self.shape[0]
        s.kv_cap_text, s.win_cap = self.kv_cap_text, self.win_cap
        s.n_lm_rows = self.lm_rows.shape[0]
        s.pairs_text_self, s.pairs_i2t = float(self.pairs_text_self), float(self.pairs_i2t)
        s.n_shared_rows = int(getattr(self, "n_shared_rows", 0))
        s.no_cls_rows = int(bool(getattr(self, "scores_only", False)))
        s.n_lm_unique = self.lm_urows.shape[0] if self.lm_urows.shape[0] < self.lm_rows.shape[0] else 0
        if getattr(self, "unit_image", None) is None:
            s.d_unit_image, s.n_images = None, 0      # this (specification) packer stores one feature block per unit
        else:
            s.n_images = int(self.image_feat.shape[0])
        return s


def _row_range_of(a: np.ndarray):
    """(base, first_row) when ``a`` is a row range of a C-contiguous 2-D array with the same row length, else (a itself, 0)."""
    b = a.base
    if isinstance(b, np.ndarray) and b.ndim == 2 and a.ndim == 2 and b.shape[1] == a.shape[1] and b.dtype == a.dtype \
            and b.flags.c_contiguous and a.flags.c_contiguous:
        off = a.__array_interface__["data"][0] - b.__array_interface__["data"][0]
        row_bytes = b.shape[1] * b.itemsize
        if off >= 0 and off % row_bytes == 0 and off // row_bytes + a.shape[0] <= b.shape[0]:
            return b, off // row_bytes
    return np.ascontiguousarray(a), 0


def _roundup(x: int, m: int) -> int:
    return (x + m - 1) // m * m


def pack_units(units: Sequence[UnitArrays], image_feat: np.ndarray, image_loc: np.ndarray, image_mask: np.ndarray,
               R: int = R_DEFAULT, verify_shared: bool = True, scores_only: bool = False, share_first_mask: bool = True) -> PackedBatch:
    """Pack generative-mode units.  ``image_*`` hold one block per *slot* ([n_slots,R,...]); a unit's image rows are
    gathered from ``unit.image_slot`` (so the 10 rounds of an image can share one host copy)."""
    U = len(units)
    n_cls = 0 if scores_only else 1           # [CLS] rows per candidate
    a_drop = 1 if scores_only else 0          # visible-copy rows dropped from the end (A_{last-1})
    # ---- per-unit checks (the only per-unit work besides the gathers further down; everything else is vectorised over all
    # candidates / rows of the batch, so that packing a bench-size step — 80 units, 8 000 candidates — stays around 10 ms)
    n_u = np.asarray([len(u.desc) for u in units], np.int64)
    desc = np.concatenate([np.asarray(u.desc) for u in units]).astype(np.int64)          # [C,4]
    C_tot = int(desc.shape[0])
    ctx_u = np.asarray([int(u.desc[0, 1]) for u in units], np.int64)
    unit_of = np.repeat(np.arange(U), n_u)                                               # [C] unit of every candidate
    if (desc[:, 0] != 0).any():
        raise ValueError("prefix sharing applies to generative-mode sequences only")
    if (desc[:, 1] != ctx_u[unit_of]).any() or (ctx_u < 2).any():
        raise ValueError("all candidates of a unit must share one context of at least one token")
    last, L = desc[:, 3], desc[:, 2]
    if verify_shared:
        for ui, u in enumerate(units):
            ctx = int(ctx_u[ui])
            if n_u[ui] > 1:
                same = (u.tokens[:, 1:ctx] == u.tokens[0, 1:ctx]).all() and (u.segments[:, 1:ctx] == u.segments[0, 1:ctx]).all() \
                    and (u.positions[:, 1:ctx] == u.positions[0, 1:ctx]).all()
                if not same:
                    raise ValueError("candidates of a unit differ in their context rows: cannot share the prefix")
    # ---- source arrays: consecutive units that are row ranges of the same arrays (the rounds of an image, as the reference's
    # loader stores them) are gathered from with ONE numpy call per field; a stand-alone unit is a group of its own
    S = units[0].tokens.shape[1]
    first_c = np.concatenate([[0], np.cumsum(n_u)])[:-1]          # first candidate of every unit
    src = [[_row_range_of(a) for a in (u.tokens, u.segments, u.positions, u.labels)] for u in units]
    row0_u = np.asarray([[r0 for _, r0 in fields] for fields in src], np.int64)            # [U,4] first row inside the base
    groups, g0 = [], 0                                                                    # (first unit, end unit)
    for ui in range(1, U + 1):
        if ui == U or any(src[ui][f][0] is not src[g0][f][0] for f in range(4)):
            groups.append((g0, ui))
            g0 = ui
    c_end = np.concatenate([first_c, [C_tot]])
    local_c = np.arange(C_tot) - first_c[unit_of]                 # candidate index inside its unit
    # B_0 (the first masked position) sees the context and itself only; if its token / segment / position agree across the
    # candidates its whole row is the same for all of them: one row per unit (scores-only layout: nothing but [CLS] attends it)
    b0_shared = np.zeros(U, np.int64)
    if scores_only and share_first_mask:
        ok = np.ones(C_tot, bool)
        for g_lo, g_hi in groups:
            c0, c1 = int(c_end[g_lo]), int(c_end[g_hi])
            for f in range(3):
                v = np.take(src[g_lo][f][0].reshape(-1), (row0_u[unit_of[c0:c1], f] + local_c[c0:c1]) * S + L[c0:c1])
                ok[c0:c1] &= v == v[first_c[unit_of[c0:c1]] - c0]
        b0_shared = (np.logical_and.reduceat(ok, first_c) & (n_u > 1)).astype(np.int64)
    # ---- row layout: [all units' context rows | per unit: its shared B_0 row (if any), candidate 0's rows, candidate 1's rows, ...]
    sh_len = ctx_u - 1
    sh_start = np.concatenate([[0], np.cumsum(sh_len)])
    n_shared = int(sh_start[-1])
    b_drop = b0_shared[unit_of]                                   # [C] 1: this candidate's B_0 lives in the unit's shared row
    rep = n_cls - a_drop - b_drop + 2 * last                      # [C] own rows: [CLS] (n_cls), A_0..A_{na-1}, B_{b_drop}..B_{last-1}
    rep_excl = np.cumsum(rep) - rep                               # own rows before candidate c (whole batch)
    unit_own = np.add.reduceat(rep, first_c)                      # own rows of every unit (n_u >= 1)
    unit_rows = b0_shared + unit_own
    unit_base = n_shared + np.concatenate([[0], np.cumsum(unit_rows)])
    M = int(unit_base[-1])
    cs = unit_base[unit_of] + b_drop + (rep_excl - rep_excl[first_c][unit_of])     # [C] absolute first row of each candidate
    n_own = int(rep.sum())
    owner = np.repeat(np.arange(C_tot), rep)                      # [n_own] candidate of every own row
    idx = np.arange(n_own) - np.repeat(rep_excl, rep)             # row index inside the candidate
    dst = np.repeat(cs, rep) + idx                                # absolute packed row
    last_r, ctx_r = last[owner], ctx_u[unit_of][owner]
    na_r = last_r - a_drop
    is_cls, is_b = idx < n_cls, idx >= n_cls + na_r
    is_a = ~is_cls & ~is_b
    k = np.where(is_b, idx - n_cls - na_r + b_drop[owner], idx - n_cls)             # index inside the A / B copy
    src_col = np.where(is_cls, 0, np.where(is_a, ctx_r + k, ctx_r + last_r + k))    # dense column this packed row comes from
    a0 = np.repeat(cs, rep) + n_cls                               # first A row of the candidate
    row_iv = np.zeros((M, 4), np.int32)
    row_iv[:, 2] = -1
    row_iv[dst, 0] = np.where(is_cls, a0 - n_cls, a0)
    row_iv[dst, 1] = np.where(is_cls, a0 - n_cls + rep[owner], np.where(is_a, a0 + k + 1, a0 + k))
    row_iv[dst, 2] = np.where(is_b, dst, -1)
    own_keys = np.where(is_cls, rep[owner], k + 1)                # own-candidate keys incl. self
    pairs_ts = int((sh_len ** 2).sum() + ((ctx_r - 1) + own_keys).sum() + (b0_shared * ctx_u).sum())
    pairs_i2t = int(R * sh_len.sum())
    b0_rows = unit_base[:-1][b0_shared == 1]
    row_iv[b0_rows, 0] = row_iv[b0_rows, 1] = row_iv[b0_rows, 2] = b0_rows          # no own-candidate keys besides itself
    # ---- labelled rows of each candidate in order B_0..B_{last-1}
    lm_off = np.concatenate([[0], np.cumsum(last)])
    owner_l = np.repeat(np.arange(C_tot), last)
    k_l = np.arange(int(lm_off[-1])) - np.repeat(lm_off[:-1], last)
    bd_l = b_drop[owner_l]
    lm_rows_all = np.where((k_l == 0) & (bd_l == 1), unit_base[unit_of[owner_l]],
                           cs[owner_l] + n_cls + (last[owner_l] - a_drop) + k_l - bd_l)
    lab_col = L[owner_l] + k_l
    # ---- gathers from the units' dense arrays (per unit: its context rows, its own rows, its labels)
    ids = np.zeros(M, np.int32)
    segs = np.zeros(M, np.int32)
    pos = np.zeros(M, np.int32)
    lm_labels = np.empty(int(lm_off[-1]), np.int32)
    own_off = np.concatenate([[0], np.cumsum(unit_own)])
    lm_unit_off = lm_off[c_end]
    u_own, u_lm = unit_of[owner], unit_of[owner_l]
    ctx_unit = np.repeat(np.arange(U), sh_len)                    # unit of every context row
    ctx_col = 1 + np.arange(n_shared) - sh_start[:-1][ctx_unit]   # its dense column (candidate 0 of the unit)
    own_dst = unit_base[:-1] + b0_shared                          # the unit's own rows are the contiguous range [own_dst, + unit_own)
    for g_lo, g_hi in groups:
        o0, o1 = int(own_off[g_lo]), int(own_off[g_hi])
        s0, s1 = int(sh_start[g_lo]), int(sh_start[g_hi])
        l0_, l1_ = int(lm_unit_off[g_lo]), int(lm_unit_off[g_hi])
        b0u = np.flatnonzero(b0_shared[g_lo:g_hi]) + g_lo
        for f, dst_arr in enumerate((ids, segs, pos)):
            base = src[g_lo][f][0].reshape(-1)
            dst_arr[s0:s1] = np.take(base, row0_u[ctx_unit[s0:s1], f] * S + ctx_col[s0:s1])
            dst_arr[dst[o0:o1]] = np.take(base, (row0_u[u_own[o0:o1], f] + local_c[owner[o0:o1]]) * S + src_col[o0:o1])
            dst_arr[unit_base[b0u]] = np.take(base, row0_u[b0u, f] * S + L[first_c[b0u]])
        lm_labels[l0_:l1_] = np.take(src[g_lo][3][0].reshape(-1), (row0_u[u_lm[l0_:l1_], 3] + local_c[owner_l[l0_:l1_]]) * S + lab_col[l0_:l1_])
    if (lm_labels < 0).any():
        raise ValueError("a masked-copy position carries no label")
    # distinct labelled rows in ascending order (the shared B_0 rows appear once) and every entry's index among them
    is_lm = np.zeros(M, bool)
    is_lm[lm_rows_all] = True
    lm_urows = np.flatnonzero(is_lm)
    lm_uidx = (np.cumsum(is_lm) - 1)[lm_rows_all]
    # ---- attention jobs (8 int32 each: q_start, q_len, kv_start, kv_len, win, mask_row, 0, 0)
    zu, uu = np.zeros(U, np.int64), np.arange(U)
    s0_u, q0_u = sh_start[:-1], unit_base[:-1]
    job = lambda *cols: np.stack([np.asarray(c, np.int64) + zu for c in cols], 1)
    jobs_ctx = job(s0_u, sh_len, s0_u, sh_len, 0, -1, 0, 0)
    jobs_cand = job(q0_u, unit_rows, s0_u, sh_len, 1, -1, 0, 0)
    jobs_t2i = np.stack([job(s0_u, sh_len, uu * R, R, 0, uu, 0, 0), job(q0_u, unit_rows, uu * R, R, 0, uu, 0, 0)], 1).reshape(2 * U, 8)
    jobs_i2t = job(uu * R, R, s0_u, sh_len, 0, -1, 0, 0)
    jobs_img = job(uu * R, R, uu * R, R, 0, uu, 0, 0)
    max_cand_q = max(1, int(unit_rows.max()))
    slots = np.asarray([u.image_slot for u in units])
    t = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a, dtype=dt))
    max_rows_per_cand = max(1, int(rep.max()))
    max_sh = int(sh_len.max())
    return PackedBatch(
        n_units=U, n_cands=C_tot, n_text_rows=M, n_shared_rows=n_shared, scores_only=bool(scores_only), n_b0_shared=int(b0_shared.sum()),
        input_ids=t(ids, np.int32), token_type_ids=t(segs, np.int32), position_ids=t(pos, np.int32), row_iv=t(row_iv, np.int32),
        jobs_text_self=t(np.concatenate([jobs_ctx, jobs_cand]), np.int32), n_jobs_text_ctx=U, cand_halo=max_rows_per_cand - 1,
        jobs_t2i=t(jobs_t2i, np.int32), jobs_i2t=t(jobs_i2t, np.int32), jobs_img_self=t(jobs_img, np.int32),
        lm_rows=t(lm_rows_all, np.int32), lm_labels=t(lm_labels, np.int32), lm_urows=t(lm_urows, np.int32), lm_uidx=t(lm_uidx, np.int32),
        cand_lm_off=t(lm_off, np.int32), cand_cls_row=t(cs if n_cls else np.full(C_tot, -1), np.int32), cand_img_row=t(unit_of * R, np.int32),
        image_feat=t(image_feat[slots], np.float32), image_loc=t(image_loc[slots], np.float32), image_mask=t(image_mask[slots], np.float32),
        max_q_text_self=max(max_cand_q, max_sh), max_q_t2i=max(max_cand_q, max_sh),
        kv_cap_text=_roundup(max_sh, 64), win_cap=_roundup(Q_TILE + 2 * (max_rows_per_cand - 1), 64),
        pairs_text_self=pairs_ts, pairs_i2t=pairs_i2t, n_dense_rows=C_tot * units[0].tokens.shape[1])


def units_from_rounds(rounds, image_slots: Optional[List[int]] = None) -> List[UnitArrays]:
    """``unimm_b200.synthetic.Round`` objects -> units (one image slot per round unless given)."""
    return [UnitArrays(r.tokens, r.segments, r.positions, r.labels, r.desc, (image_slots[i] if image_slots else i))
            for i, r in enumerate(rounds)]


def units_from_flat(tokens, segments, positions, labels, desc, unit_index) -> List[UnitArrays]:
    """Reference-format flat batch ([B,S] tensors + descriptors) grouped by ``unit_index`` [B] (non-decreasing)."""
    tok, seg, pos, lab, d, ui = (np.asarray(x) for x in (tokens, segments, positions, labels, desc, unit_index))
    if ui.size and (np.diff(ui) < 0).any():
        raise ValueError("unit_index must be non-decreasing")
    cuts = np.concatenate([[0], np.flatnonzero(np.diff(ui)) + 1, [ui.size]])
    # row ranges (views) of the flat arrays: the packer gathers from the flat arrays directly
    return [UnitArrays(tok[a:b], seg[a:b], pos[a:b], lab[a:b], d[a:b], int(ui[a])) for a, b in zip(cuts[:-1], cuts[1:]) if b > a]
