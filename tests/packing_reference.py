"""Test oracle for the packer: the straightforward per-unit implementation of ``unimm_b200.packing.pack_units`` (one Python
loop iteration per unit, one numpy statement per quantity).  The product packer is vectorised over the whole batch; this version is
kept so that ``tests/test_packing_cpu.py`` can require the two to produce identical batches, tensor for tensor."""
from typing import Sequence

import numpy as np
import torch

from unimm_b200.packing import Q_TILE, R_DEFAULT, PackedBatch, UnitArrays, _roundup


def pack_units_loop(units: Sequence[UnitArrays], image_feat: np.ndarray, image_loc: np.ndarray, image_mask: np.ndarray,
               R: int = R_DEFAULT, verify_shared: bool = True, scores_only: bool = False, share_first_mask: bool = True) -> PackedBatch:
    """Pack generative-mode units.  ``image_*`` hold one block per *slot* ([n_slots,R,...]); a unit's image rows are
    gathered from ``unit.image_slot`` (so the 10 rounds of an image can share one host copy)."""
    U = len(units)
    n_cls = 0 if scores_only else 1           # [CLS] rows per candidate
    a_drop = 1 if scores_only else 0          # visible-copy rows dropped from the end (A_{last-1})
    sh_len, n_cand, cand_rows, b0_shared = [], [], [], []
    for u in units:
        d = u.desc
        if (d[:, 0] != 0).any():
            raise ValueError("prefix sharing applies to generative-mode sequences only")
        ctx = int(d[0, 1])
        if (d[:, 1] != ctx).any() or ctx < 2:
            raise ValueError("all candidates of a unit must share one context of at least one token")
        if verify_shared and len(d) > 1:
            same = (u.tokens[:, 1:ctx] == u.tokens[0, 1:ctx]).all() and (u.segments[:, 1:ctx] == u.segments[0, 1:ctx]).all() \
                and (u.positions[:, 1:ctx] == u.positions[0, 1:ctx]).all()
            if not same:
                raise ValueError("candidates of a unit differ in their context rows: cannot share the prefix")
        # B_0 (the first masked position) sees the context and itself only; if its token / segment / position agree across the
        # candidates its whole row is the same for all of them: one row per unit (scores-only layout: nothing but [CLS] attends it)
        Lc = d[:, 2].astype(np.int64)
        ar = np.arange(len(d))
        b0_same = scores_only and share_first_mask and len(d) > 1 and (u.tokens[ar, Lc] == u.tokens[0, Lc[0]]).all() \
            and (u.segments[ar, Lc] == u.segments[0, Lc[0]]).all() and (u.positions[ar, Lc] == u.positions[0, Lc[0]]).all()
        b0_shared.append(bool(b0_same))
        sh_len.append(ctx - 1)
        n_cand.append(len(d))
        cand_rows.append(n_cls - a_drop - (1 if b0_same else 0) + 2 * d[:, 3].astype(np.int64))
    sh_start = np.concatenate([[0], np.cumsum(sh_len)])
    n_shared = int(sh_start[-1])
    all_rows = np.concatenate(cand_rows)
    C_tot = int(all_rows.shape[0])
    # candidate block of unit ui: [its shared B_0 row, if any | candidate 0's rows | candidate 1's rows | ...]
    unit_rows = np.asarray([int(b) + int(r.sum()) for b, r in zip(b0_shared, cand_rows)], np.int64)
    unit_base = n_shared + np.concatenate([[0], np.cumsum(unit_rows)])
    M = int(unit_base[-1])
    ids = np.zeros(M, np.int32)
    segs = np.zeros(M, np.int32)
    pos = np.zeros(M, np.int32)
    row_iv = np.zeros((M, 4), np.int32)
    row_iv[:, 2] = -1
    lm_rows, lm_labels, cand_lm_off = [], [], [0]
    cls_row = np.zeros(C_tot, np.int32)
    img_row = np.zeros(C_tot, np.int32)
    jobs_ctx, jobs_cand, jobs_t2i, jobs_i2t, jobs_img = [], [], [], [], []
    pairs_ts = pairs_i2t = 0
    ci = 0
    max_cand_q = 1
    for ui, u in enumerate(units):
        ctx = sh_len[ui] + 1
        s0 = int(sh_start[ui])
        ids[s0:s0 + ctx - 1] = u.tokens[0, 1:ctx]
        segs[s0:s0 + ctx - 1] = u.segments[0, 1:ctx]
        pos[s0:s0 + ctx - 1] = u.positions[0, 1:ctx]
        n = n_cand[ui]
        last = u.desc[:, 3].astype(np.int64)
        L = u.desc[:, 2].astype(np.int64)
        b_drop = 1 if b0_shared[ui] else 0
        rep = cand_rows[ui]                                       # rows of each candidate: [CLS] (n_cls), A_0..A_{na-1}, B_{b_drop}..B_{last-1}
        q0 = int(unit_base[ui])                                   # first row of the unit's candidate block
        rows_u = int(unit_rows[ui])
        cs = q0 + b_drop + np.concatenate([[0], np.cumsum(rep)])[:-1]     # absolute first row of each candidate
        n_own = int(rep.sum())
        owner = np.repeat(np.arange(n), rep)
        idx = np.arange(n_own) - np.repeat(cs - cs[0], rep)
        s_abs = np.repeat(cs, rep)
        last_r = np.repeat(last, rep)
        na_r = last_r - a_drop
        is_cls, is_b = idx < n_cls, idx >= n_cls + na_r
        is_a = ~is_cls & ~is_b
        k = np.where(is_b, idx - n_cls - na_r + b_drop, idx - n_cls)      # index inside the A / B copy
        src_col = np.where(is_cls, 0, np.where(is_a, ctx + k, ctx + last_r + k))   # dense column this packed row comes from
        dst = int(cs[0]) + np.arange(n_own)
        ids[dst] = u.tokens[owner, src_col]
        segs[dst] = u.segments[owner, src_col]
        pos[dst] = u.positions[owner, src_col]
        a0 = s_abs + n_cls                                        # first A row of the candidate
        lo = np.where(is_cls, s_abs, a0)
        hi = np.where(is_cls, s_abs + rep[owner], np.where(is_a, a0 + k + 1, a0 + k))
        row_iv[dst, 0], row_iv[dst, 1] = lo, hi
        row_iv[dst, 2] = np.where(is_b, dst, -1)
        own_keys = np.where(is_cls, rep[owner], k + 1)            # own-candidate keys incl. self
        pairs_ts += (ctx - 1) ** 2 + int(((ctx - 1) + own_keys).sum())
        pairs_i2t += R * (ctx - 1)
        # labelled rows of each candidate in order B_0..B_{last-1}
        off = np.concatenate([[0], np.cumsum(last)])
        owner_l = np.repeat(np.arange(n), last)
        k_l = np.arange(int(off[-1])) - np.repeat(off[:-1], last)
        lm_u = np.empty(int(off[-1]), np.int64)
        if b_drop:
            ids[q0], segs[q0], pos[q0] = u.tokens[0, L[0]], u.segments[0, L[0]], u.positions[0, L[0]]
            row_iv[q0] = (q0, q0, q0, 0)                          # no own-candidate keys besides itself
            pairs_ts += ctx
            lm_u[k_l == 0] = q0
            lm_u[k_l > 0] = dst[is_b]
        else:
            lm_u[:] = dst[is_b]
        lm_rows.append(lm_u)
        lm_labels.append(u.labels[owner_l, L[owner_l] + k_l])
        cand_lm_off.extend((cand_lm_off[-1] + np.cumsum(last)).tolist())
        cls_row[ci:ci + n] = cs if n_cls else -1
        img_row[ci:ci + n] = ui * R
        jobs_ctx.append((s0, ctx - 1, s0, ctx - 1, 0, -1, 0, 0))
        jobs_cand.append((q0, rows_u, s0, ctx - 1, 1, -1, 0, 0))
        jobs_t2i.append((s0, ctx - 1, ui * R, R, 0, ui, 0, 0))
        jobs_t2i.append((q0, rows_u, ui * R, R, 0, ui, 0, 0))
        jobs_i2t.append((ui * R, R, s0, ctx - 1, 0, -1, 0, 0))
        jobs_img.append((ui * R, R, ui * R, R, 0, ui, 0, 0))
        max_cand_q = max(max_cand_q, rows_u)
        ci += n
    lm_rows_all = np.concatenate(lm_rows)
    lm_urows, lm_uidx = np.unique(lm_rows_all, return_inverse=True)     # distinct labelled rows (the shared B_0 rows appear once)
    lm_labels = np.concatenate(lm_labels).astype(np.int32)
    if (lm_labels < 0).any():
        raise ValueError("a masked-copy position carries no label")
    slots = np.asarray([u.image_slot for u in units])
    t = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a, dtype=dt))
    max_rows_per_cand = max(1, int(all_rows.max()))
    return PackedBatch(
        n_units=U, n_cands=C_tot, n_text_rows=M, n_shared_rows=n_shared, scores_only=bool(scores_only), n_b0_shared=int(sum(b0_shared)),
        input_ids=t(ids, np.int32), token_type_ids=t(segs, np.int32), position_ids=t(pos, np.int32), row_iv=t(row_iv, np.int32),
        jobs_text_self=t(np.asarray(jobs_ctx + jobs_cand), np.int32), n_jobs_text_ctx=len(jobs_ctx), cand_halo=max_rows_per_cand - 1, jobs_t2i=t(np.asarray(jobs_t2i), np.int32),
        jobs_i2t=t(np.asarray(jobs_i2t), np.int32), jobs_img_self=t(np.asarray(jobs_img), np.int32),
        lm_rows=t(lm_rows_all, np.int32), lm_labels=t(lm_labels, np.int32), lm_urows=t(lm_urows, np.int32), lm_uidx=t(lm_uidx, np.int32),
        cand_lm_off=t(np.asarray(cand_lm_off), np.int32), cand_cls_row=t(cls_row, np.int32), cand_img_row=t(img_row, np.int32),
        image_feat=t(image_feat[slots], np.float32), image_loc=t(image_loc[slots], np.float32), image_mask=t(image_mask[slots], np.float32),
        max_q_text_self=max(max_cand_q, max(sh_len)), max_q_t2i=max(max_cand_q, max(sh_len)),
        kv_cap_text=_roundup(max(sh_len), 64), win_cap=_roundup(Q_TILE + 2 * (max_rows_per_cand - 1), 64),
        pairs_text_self=pairs_ts, pairs_i2t=pairs_i2t, n_dense_rows=C_tot * units[0].tokens.shape[1])


