"""CPU: the C++ packer (csrc/packer.cu behind unimm_packer_*, the one the sweep and the bench call) against the numpy
specification (unimm_b200/packing.py) array for array, against the dense reference masks row for row — truncated sequences
(utils/data_utils.py:205-209, :237-244) included — and its error behaviour.  Host code only: no GPU needed."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from unimm_b200 import synthetic as syn
from unimm_b200._lib import UnimmError
from unimm_b200.descriptors import dense_text_mask, descriptors_from_masks
from unimm_b200.flat_packer import FlatPacker, ImageArrays, view_to_batch
from unimm_b200.packing import pack_units, units_from_rounds

R, F = 37, 2048
INT_KEYS = ("input_ids", "token_type_ids", "position_ids", "row_iv", "jobs_text_self", "jobs_i2t", "lm_rows", "lm_labels", "cand_lm_off",
            "cand_cls_row", "cand_img_row", "lm_urows", "lm_uidx")
SCALARS = ("n_units", "n_cands", "n_text_rows", "n_shared_rows", "max_q_text_self", "max_q_t2i", "cand_halo", "kv_cap_text", "win_cap",
           "n_jobs_text_ctx")


def _images(n_images, rounds_of, n_cand):
    imgs, rounds, slots, blocks = [], [], [], []
    for i in range(n_images):
        (f, l, m), rs = syn.synth_dialog_rounds(40 + i, rounds=rounds_of(i), n_candidates=n_cand)
        imgs.append(ImageArrays.from_rounds(rs, f, l, m))
        rounds += rs
        slots += [i] * len(rs)
        blocks.append((f, l, m))
    return imgs, rounds, slots, blocks


@pytest.mark.parametrize("scores_only,share", [(True, True), (True, False), (False, False)])
def test_cpp_packer_equals_numpy_specification(scores_only, share):
    imgs, rounds, slots, blocks = _images(3, lambda i: (1, 4, 10) if i != 1 else (2, 7), 23)
    pk = FlatPacker(pinned=False, threads=3)
    v = pk.pack(imgs, scores_only=scores_only, share_first_mask=share)
    feat, loc, mask = (np.stack([b[k] for b in blocks]) for k in range(3))
    pb = pack_units(units_from_rounds(rounds, slots), feat, loc, mask, scores_only=scores_only, share_first_mask=share)
    a, s = v.arrays(R, F), v.struct
    for k in INT_KEYS:
        ref = getattr(pb, k).numpy()
        if k in ("lm_urows", "lm_uidx") and not s.n_lm_unique:
            continue                                          # nothing shared: the C struct carries no distinct-row lists
        assert a[k].shape == ref.shape and np.array_equal(a[k], ref), k
    for k in SCALARS:
        assert getattr(s, k) == getattr(pb, k), k
    assert s.pairs_text_self == pb.pairs_text_self and s.pairs_i2t == pb.pairs_i2t
    assert s.no_cls_rows == int(scores_only)
    # one feature block per IMAGE; the jobs' mask row and unit_image name it
    slots = np.asarray(slots)
    assert np.array_equal(a["unit_image"], slots) and s.n_images == 3
    for k in range(3):
        assert np.array_equal(a[("image_feat", "image_loc", "image_mask")[k]], (feat, loc, mask)[k])
    for name in ("jobs_t2i", "jobs_img_self"):
        ref = getattr(pb, name).numpy().copy()
        ref[:, 5] = slots[ref[:, 5]]
        assert np.array_equal(a[name], ref), name
    # the torch-tensor copy used for device-resident runs describes the same batch
    b2 = view_to_batch(v, R, F)
    c2 = b2.c_struct()
    assert c2.n_images == 3 and c2.n_lm_unique == s.n_lm_unique and c2.n_text_rows == s.n_text_rows
    pk.close()


def test_descriptors_derived_from_position_ids():
    imgs, rounds, _, _ = _images(2, lambda i: (1, 6, 10), 17)
    bare = [ImageArrays(im.tokens, im.segments, im.positions, im.labels, im.feat, im.loc, im.mask, units=im.units) for im in imgs]
    pk = FlatPacker(pinned=False)
    v1 = pk.pack(imgs)
    a1 = v1.arrays(R, F)
    v2 = pk.pack(bare)
    assert np.array_equal(pk.desc(), np.concatenate([r.desc for r in rounds]))
    a2 = v2.arrays(R, F)
    for k in a1:
        assert np.array_equal(a1[k], a2[k]), k


def _check_against_dense(v, tokens, segments, positions, labels, desc, unit_ranges, scores_only, S=256):
    """Every packed candidate row stands for one dense position: same token / segment / position, and its key set (context +
    own-candidate interval + self) equals the dense mask row restricted to [0, S)."""
    a, s = v.arrays(R, F), v.struct
    iv, jobs, off, lm_rows, lab = a["row_iv"], a["jobs_text_self"], a["cand_lm_off"], a["lm_rows"], a["lm_labels"]
    dm_all = dense_text_mask(torch.from_numpy(desc), S).numpy()
    c = 0
    for ui, (r0, n) in enumerate(unit_ranges):
        ctx = int(desc[r0, 1])
        sh, cj = jobs[ui], jobs[s.n_jobs_text_ctx + ui]
        assert sh[1] == min(ctx, S) - 1 and np.array_equal(a["input_ids"][sh[0]:sh[0] + sh[1]], tokens[r0, 1:1 + sh[1]])
        row = int(cj[0])
        b0_row = -1
        if iv[row, 2] == row and iv[row, 0] == row and iv[row, 1] == row and scores_only:   # the unit's shared B_0 row
            b0_row, row = row, row + 1
        for j in range(r0, r0 + n):
            L, last = int(desc[j, 2]), int(desc[j, 3])
            nb = max(0, min(last, S - L))
            na = min(last, S - ctx)
            if scores_only:
                na = min(na, max(nb - 1, 0))
            b0 = 1 if (b0_row >= 0 and nb > 0) else 0
            cols = ([] if scores_only else [0]) + [ctx + k for k in range(na)] + [L + k for k in range(b0, nb)]
            first = row
            for col in cols:
                allowed = set(range(1, min(ctx, S)))
                lo, hi, sf = iv[row, :3]
                for q in list(range(lo, hi)) + ([sf] if sf >= 0 else []):
                    assert 0 <= q - first < len(cols), "a row may only see rows of its own candidate"
                    allowed.add(cols[q - first])
                want = set(np.nonzero(dm_all[j, col])[0].tolist())
                if scores_only:
                    want -= {0}                  # [CLS] sees everything but nothing kept sees [CLS]; not packed
                assert allowed == want, (ui, j, col)
                assert a["input_ids"][row] == tokens[j, col] and a["position_ids"][row] == positions[j, col]
                assert a["token_type_ids"][row] == segments[j, col]
                row += 1
            # labelled rows: every masked-copy position that exists, in order, with the reference's labels
            assert off[c + 1] - off[c] == nb == int((labels[j] != -1).sum())
            assert np.array_equal(lab[off[c]:off[c + 1]], labels[j, L:L + nb])
            want_rows = ([b0_row] if b0 else []) + list(range(row - (nb - b0), row))
            assert lm_rows[off[c]:off[c + 1]].tolist() == want_rows
            c += 1
        assert row == cj[0] + cj[1]
    assert c == s.n_cands


@pytest.mark.parametrize("scores_only", [True, False])
@pytest.mark.parametrize("derive", [False, True])
def test_truncated_sequences_keep_the_rows_that_exist(scores_only, derive):
    """The reference's own encoder output for a 247-position context (tests/golden/gen10_truncated.npz): answers whose masked
    copy is cut by S = 256, has no room at all, or whose visible copy is cut too."""
    g, b = load_golden("gen10_truncated")
    desc = descriptors_from_masks(b["txt_attention_mask"], b["co_attention_mask"]).numpy()
    assert (desc[:, 2] + desc[:, 3] > 256).sum() == 6 and (desc[:, 2] >= 256).sum() == 2
    tok, seg, pos, lab = (b[k].numpy() for k in ("tokens", "segments", "positions", "mask"))
    im = ImageArrays(tok, seg, pos, lab, g["image_feat"], g["image_loc"], g["image_mask"], desc=None if derive else desc, units=[(0, 10)])
    pk = FlatPacker(pinned=False)
    v = pk.pack([im], scores_only=scores_only)
    d_used = pk.desc()
    if derive:           # candidates 5 and 6 have no masked copy inside S: L = S stands for "beyond", ctx comes from the siblings
        assert np.array_equal(d_used[:, 1], desc[:, 1]) and np.array_equal(np.minimum(d_used[:, 2], 256), np.minimum(desc[:, 2], 256))
    _check_against_dense(v, tok, seg, pos, lab, desc, [(0, 10)], scores_only)
    assert v.n_lm_rows == int((lab != -1).sum())


def test_untruncated_batch_against_dense_masks():
    imgs, rounds, _, _ = _images(1, lambda i: (1, 5, 10), 31)
    im = imgs[0]
    pk = FlatPacker(pinned=False)
    for so in (True, False):
        v = pk.pack([im], scores_only=so)
        _check_against_dense(v, im.tokens, im.segments, im.positions, im.labels, im.desc, im.units, so)


def test_packer_errors():
    imgs, rounds, _, _ = _images(1, lambda i: (3,), 5)
    im = imgs[0]
    pk = FlatPacker(pinned=False)
    bad = ImageArrays(im.tokens.copy(), im.segments, im.positions, im.labels, im.feat, im.loc, im.mask, desc=im.desc, units=im.units)
    bad.tokens[3, 7] += 1                                   # candidate 3's context differs from candidate 0's
    with pytest.raises(UnimmError, match="differ in their context"):
        pk.pack([bad])
    pk.pack([bad], verify_shared=False)                      # the check is what finds it
    bad = ImageArrays(im.tokens, im.segments, im.positions, im.labels.copy(), im.feat, im.loc, im.mask, desc=im.desc, units=im.units)
    bad.labels[2, int(im.desc[2, 2])] = -1
    with pytest.raises(UnimmError, match="carries no label"):
        pk.pack([bad])
    d = im.desc.copy()
    d[1, 0] = 1
    with pytest.raises(UnimmError, match="generative-mode"):
        pk.pack([ImageArrays(im.tokens, im.segments, im.positions, im.labels, im.feat, im.loc, im.mask, desc=d, units=im.units)])
    with pytest.raises(UnimmError, match="row range outside"):
        pk.pack([ImageArrays(im.tokens, im.segments, im.positions, im.labels, im.feat, im.loc, im.mask, desc=im.desc, units=[(2, 9)])])
    with pytest.raises(ValueError):
        pk.pack([ImageArrays(im.tokens[:, :100], im.segments[:, :100], im.positions[:, :100], im.labels[:, :100], im.feat, im.loc, im.mask)])
    v = pk.pack([im])                                        # the packer still works after failed calls
    assert v.n_cands == 5
