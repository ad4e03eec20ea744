"""Host-side packing from the reference's own data layout (``csrc/packer.cu`` behind ``unimm_packer_*``).

``val_lm.py:55-121`` holds, per image, one int64 ``[rounds*options, 256]`` tensor per field (``dataloader_visdial.py:437-457``)
and one ``[37, 2048]`` feature block.  ``FlatPacker.pack`` takes a step of such images as they are — no concatenation, no dense
masks — and produces the prefix-shared batch of ``include/unimm_b200.h`` inside pinned staging buffers that
``Engine.score_packed_host`` uploads.  Two packers double-buffer a sweep (``unimm_b200.val_sweep``).
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence

import numpy as np

from ._lib import FlatBatch, ImageBlock, PackedBatchStruct, SeqDesc, check, lib


class ImageArrays:
    """One image of a step: int64 ``[rows, S]`` arrays + optional int32 ``[rows, 4]`` descriptors + its feature block.

    ``units`` lists the (first_row, n_rows) range of every unit = (image, round) inside the arrays; default: ``rows_per_unit``
    consecutive rows each."""

    def __init__(self, tokens, segments, positions, labels, feat, loc, mask, desc=None, units: Optional[Sequence] = None,
                 rows_per_unit: int = 100):
        as64 = lambda a: np.ascontiguousarray(a, dtype=np.int64)
        self.tokens, self.segments, self.positions, self.labels = as64(tokens), as64(segments), as64(positions), as64(labels)
        self.desc = None if desc is None else np.ascontiguousarray(desc, dtype=np.int32)
        self.feat = np.ascontiguousarray(feat, dtype=np.float32)
        self.loc = np.ascontiguousarray(loc, dtype=np.float32)
        self.mask = np.ascontiguousarray(mask, dtype=np.float32)
        rows = self.tokens.shape[0]
        if units is None:
            units = [(r, min(rows_per_unit, rows - r)) for r in range(0, rows, rows_per_unit)]
        self.units = [(int(a), int(n)) for a, n in units]

    @staticmethod
    def from_rounds(rounds, feat, loc, mask, with_desc: bool = True) -> "ImageArrays":
        """Rounds that are row ranges of one array per field (``synthetic.synth_dialog_rounds``) are taken as views; others are
        concatenated once."""
        from .packing import _row_range_of
        bases = [[_row_range_of(getattr(r, f)) for f in ("tokens", "segments", "positions", "labels")] for r in rounds]
        same = all(bases[i][f][0] is bases[0][f][0] for i in range(len(rounds)) for f in range(4)) and \
            all(b[0][1] == b[1][1] == b[2][1] == b[3][1] for b in bases)
        if same and bases[0][0][0].dtype == np.int64:
            arrs = [bases[0][f][0] for f in range(4)]
            units = [(b[0][1], len(r.tokens)) for b, r in zip(bases, rounds)]
        else:
            arrs = [np.concatenate([getattr(r, f) for r in rounds]) for f in ("tokens", "segments", "positions", "labels")]
            cuts = np.concatenate([[0], np.cumsum([len(r.tokens) for r in rounds])])
            units = [(int(cuts[i]), len(r.tokens)) for i, r in enumerate(rounds)]
        desc = None
        if with_desc:
            desc = np.zeros((arrs[0].shape[0], 4), np.int32)
            for (a, n), r in zip(units, rounds):
                desc[a:a + n] = r.desc
        return ImageArrays(arrs[0], arrs[1], arrs[2], arrs[3], feat, loc, mask, desc=desc, units=units)


class PackedView:
    """The batch a ``FlatPacker`` holds after ``pack``: counts + the C struct with HOST pointers (valid until the next pack)."""

    is_host = True

    def __init__(self, struct: PackedBatchStruct, h2d_bytes: int):
        self.struct = struct
        self.n_units, self.n_cands, self.n_text_rows = struct.n_units, struct.n_cands, struct.n_text_rows
        self.n_lm_rows, self.n_shared_rows = struct.n_lm_rows, struct.n_shared_rows
        self._bytes = h2d_bytes

    def c_struct(self) -> PackedBatchStruct:
        return self.struct

    def bytes(self) -> int:
        return self._bytes

    def tensors(self):
        return {}

    def array(self, field: str, count: int, dtype=np.int32) -> np.ndarray:
        """Copy of one of the struct's arrays (tests)."""
        addr = getattr(self.struct, "d_" + field)
        if not addr or count == 0:
            return np.zeros(0, dtype)
        ct = C.c_int32 if dtype == np.int32 else C.c_float
        return np.ctypeslib.as_array(C.cast(addr, C.POINTER(ct)), shape=(count,)).copy()

    def arrays(self, R: int, F: int) -> dict:
        s = self.struct
        M, U, Cn, n = s.n_text_rows, s.n_units, s.n_cands, s.n_lm_rows
        NI = s.n_images if s.d_unit_image else U
        nu = s.n_lm_unique
        return {
            "input_ids": self.array("input_ids", M), "token_type_ids": self.array("token_type_ids", M),
            "position_ids": self.array("position_ids", M), "row_iv": self.array("row_iv", 4 * M).reshape(M, 4),
            "jobs_text_self": self.array("jobs_text_self", 8 * s.n_jobs_text_self).reshape(-1, 8),
            "jobs_t2i": self.array("jobs_t2i", 8 * s.n_jobs_t2i).reshape(-1, 8), "jobs_i2t": self.array("jobs_i2t", 8 * s.n_jobs_i2t).reshape(-1, 8),
            "jobs_img_self": self.array("jobs_img_self", 8 * s.n_jobs_img_self).reshape(-1, 8),
            "lm_rows": self.array("lm_rows", n), "lm_labels": self.array("lm_labels", n), "cand_lm_off": self.array("cand_lm_off", Cn + 1),
            "cand_cls_row": self.array("cand_cls_row", Cn), "cand_img_row": self.array("cand_img_row", Cn),
            "lm_urows": self.array("lm_urows", nu), "lm_uidx": self.array("lm_uidx", n if nu else 0),
            "unit_image": self.array("unit_image", U),
            "image_feat": self.array("image_feat", NI * R * F, np.float32).reshape(NI, R, F),
            "image_loc": self.array("image_loc", NI * R * 5, np.float32).reshape(NI, R, 5),
            "image_mask": self.array("image_mask", NI * R, np.float32).reshape(NI, R),
        }


def view_to_batch(view: PackedView, R: int, F: int):
    """Copy a packer's batch into a ``packing.PackedBatch`` of torch tensors (tests; device-resident runs via ``.to(device)``)."""
    import torch

    from .packing import PackedBatch
    a, s = view.arrays(R, F), view.struct
    t = {k: torch.from_numpy(v) for k, v in a.items()}
    if not s.n_lm_unique:                       # no row is shared: every labelled row is its own distinct row
        t["lm_urows"], t["lm_uidx"] = t["lm_rows"].clone(), torch.arange(s.n_lm_rows, dtype=torch.int32)
    return PackedBatch(
        n_units=s.n_units, n_cands=s.n_cands, n_text_rows=s.n_text_rows, n_shared_rows=s.n_shared_rows, scores_only=bool(s.no_cls_rows),
        n_jobs_text_ctx=s.n_jobs_text_ctx, cand_halo=s.cand_halo, max_q_text_self=s.max_q_text_self, max_q_t2i=s.max_q_t2i,
        kv_cap_text=s.kv_cap_text, win_cap=s.win_cap, pairs_text_self=s.pairs_text_self, pairs_i2t=s.pairs_i2t,
        n_dense_rows=s.n_cands * 256, **t)


class FlatPacker:
    def __init__(self, seq_len: int = 256, num_regions: int = 37, feature_size: int = 2048, pinned: bool = True, threads: int = 4):
        self.S, self.R, self.F, self.threads = seq_len, num_regions, feature_size, threads
        self._h = C.c_void_p()
        check(lib.unimm_packer_create(seq_len, num_regions, feature_size, int(pinned), C.byref(self._h)))
        self._keep = None

    def close(self) -> None:
        if getattr(self, "_h", None) is not None and self._h.value:
            lib.unimm_packer_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def pack(self, images: List[ImageArrays], scores_only: bool = True, share_first_mask: bool = True, verify_shared: bool = True) -> PackedView:
        S, R, F = self.S, self.R, self.F
        blocks = (ImageBlock * len(images))()
        ub, u0, un = [], [], []
        for i, im in enumerate(images):
            if im.tokens.ndim != 2 or im.tokens.shape[1] != S or not (im.tokens.shape == im.segments.shape == im.positions.shape == im.labels.shape):
                raise ValueError(f"image {i}: the four id arrays must be int64 [rows, {S}]")
            if im.feat.shape != (R, F) or im.loc.shape != (R, 5) or im.mask.shape != (R,):
                raise ValueError(f"image {i}: feature block must be [{R},{F}] / [{R},5] / [{R}]")
            if im.desc is not None and im.desc.shape != (im.tokens.shape[0], 4):
                raise ValueError(f"image {i}: desc must be int32 [rows, 4]")
            b = blocks[i]
            b.rows = im.tokens.shape[0]
            b.input_ids, b.token_type_ids, b.position_ids, b.masked_lm_labels = (a.ctypes.data for a in (im.tokens, im.segments, im.positions, im.labels))
            b.desc = im.desc.ctypes.data if im.desc is not None else None
            b.image_feat, b.image_loc, b.image_mask = im.feat.ctypes.data, im.loc.ctypes.data, im.mask.ctypes.data
            for a, n in im.units:
                ub.append(i), u0.append(a), un.append(n)
        ub, u0, un = (np.asarray(x, np.int32) for x in (ub, u0, un))
        fb = FlatBatch()
        fb.n_blocks, fb.blocks, fb.n_units = len(images), blocks, len(ub)
        fb.unit_block, fb.unit_row0, fb.unit_rows = ub.ctypes.data, u0.ctypes.data, un.ctypes.data
        fb.scores_only, fb.share_first_mask, fb.verify_shared = int(scores_only), int(share_first_mask), int(verify_shared)
        self._keep = (images, blocks, ub, u0, un)
        check(lib.unimm_packer_pack(self._h, C.byref(fb), self.threads))
        s = PackedBatchStruct()
        check(lib.unimm_packer_batch(self._h, C.byref(s)))
        n_int = 3 * s.n_text_rows + 4 * s.n_text_rows + 8 * (s.n_jobs_text_self + s.n_jobs_t2i + s.n_jobs_i2t + s.n_jobs_img_self) + \
            3 * s.n_lm_rows + s.n_lm_unique + 3 * s.n_cands + 1 + s.n_units
        return PackedView(s, 4 * (n_int + s.n_images * R * (F + 6)))

    def desc(self) -> np.ndarray:
        """[n_cands, 4] descriptors the last pack used (given or derived from the position ids)."""
        p, n = C.c_void_p(), C.c_int32()
        check(lib.unimm_packer_desc(self._h, C.byref(p), C.byref(n)))
        return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_int32)), shape=(n.value, 4)).copy()
