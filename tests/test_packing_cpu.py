"""CPU: the prefix-shared packer against the dense reference masks and arrays."""
import numpy as np
import torch

from conftest import load_golden
from unimm_b200 import synthetic as syn
from unimm_b200.descriptors import dense_text_mask, descriptors_from_masks
from packing_reference import pack_units_loop
from unimm_b200.packing import PackedBatch, pack_units, units_from_flat, units_from_rounds


def _check(pb, rounds, scores_only=False, shared_b0=False):
    """Every packed row carries the token / segment / position of the dense position it stands for and sees exactly the keys the
    dense mask gives that position (so every key a kept row needs is itself a kept row)."""
    iv, jobs = pb.row_iv.numpy(), pb.jobs_text_self.numpy()
    lab, off = pb.lm_labels.numpy(), pb.cand_lm_off.numpy()
    lm_rows = pb.lm_rows.numpy()
    c = 0
    shared_list = shared_b0 if isinstance(shared_b0, (list, tuple)) else [shared_b0] * len(rounds)
    for ui, r in enumerate(rounds):
        shared_b0 = shared_list[ui]
        ctx = int(r.desc[0, 1])
        sh, cj = jobs[ui], jobs[pb.n_jobs_text_ctx + ui]
        assert sh[1] == ctx - 1 and tuple(sh[:2]) == tuple(sh[2:4]) and sh[4] == 0          # context attends itself
        assert cj[3] == ctx - 1 and cj[2] == sh[0] and cj[4] == 1                           # candidates see the context + window
        assert (pb.input_ids[sh[0]:sh[0] + ctx - 1].numpy() == r.tokens[0, 1:ctx]).all()
        dm = dense_text_mask(torch.from_numpy(r.desc), 256).numpy()
        row = int(cj[0])
        b0_row = -1
        if shared_b0:                                        # one B_0 row at the head of the unit's candidate block
            b0_row, row = row, row + 1
            assert tuple(iv[b0_row, :3]) == (b0_row, b0_row, b0_row)
        for j in range(len(r.desc)):
            L, last = int(r.desc[j, 2]), int(r.desc[j, 3])
            # dense columns of the candidate's own packed rows, in packed order
            cols = ([] if scores_only else [0]) + [ctx + k for k in range(last - (1 if scores_only else 0))] + \
                [L + k for k in range(1 if shared_b0 else 0, last)]
            assert pb.cand_cls_row[c] == (-1 if scores_only else row)
            first = row
            for idx, col in enumerate(cols):
                allowed = set(range(1, ctx))
                lo, hi, sf = iv[row, :3]
                for a in list(range(lo, hi)) + ([sf] if sf >= 0 else []):
                    assert 0 <= a - first < len(cols), "a row may only see rows of its own candidate"
                    allowed.add(cols[a - first])
                assert allowed == set(np.nonzero(dm[j, col])[0].tolist()), (ui, j, idx)
                assert pb.input_ids[row] == r.tokens[j, col] and pb.position_ids[row] == r.positions[j, col]
                assert pb.token_type_ids[row] == r.segments[j, col]
                row += 1
            if shared_b0:
                # the unit's B_0 row stands for dense position L of THIS candidate too: same inputs, keys = context + itself
                assert set(np.nonzero(dm[j, L])[0].tolist()) == set(range(1, ctx)) | {L}
                assert pb.input_ids[b0_row] == r.tokens[j, L] and pb.position_ids[b0_row] == r.positions[j, L]
                assert pb.token_type_ids[b0_row] == r.segments[j, L]
                assert lm_rows[off[c]] == b0_row and (lm_rows[off[c] + 1:off[c + 1]] == np.arange(row - (last - 1), row)).all()
            else:
                assert (lm_rows[off[c]:off[c + 1]] == np.arange(row - last, row)).all()     # the labelled rows are the B rows
            assert (lab[off[c]:off[c + 1]] == r.labels[j, L:L + last]).all()
            c += 1
        assert row == cj[0] + cj[1]
    assert c == pb.n_cands and pb.win_cap % 64 == 0 and pb.kv_cap_text % 64 == 0 and pb.kv_cap_text <= 256


def test_packed_rows_reproduce_dense_masks_and_tokens():
    rng = np.random.RandomState(0)
    img = syn.synth_image(rng)
    rounds = [syn.encode_round_gen(syn.synth_context(rng, r), syn.synth_answers(rng, n)) for r, n in ((1, 7), (3, 12), (10, 9))]
    pb = pack_units(units_from_rounds(rounds, [0, 0, 0]), img[0][None], img[1][None], img[2][None])
    _check(pb, rounds)
    assert pb.n_text_rows == sum(int(r.desc[0, 1]) - 1 + int((1 + 2 * r.desc[:, 3]).sum()) for r in rounds)


def test_scores_only_packing_drops_only_rows_nothing_labelled_can_see():
    """scores_only: no [CLS], no A_{last-1}; B_0 once per unit.  _check proves that every kept row's packed key set equals its
    dense-mask key set, i.e. the kept rows are closed under "is attended by" — the dropped rows can only change the pooled NSP
    logit — and that the shared B_0 row is, for every candidate, the row the dense layout has at its position L."""
    rng = np.random.RandomState(5)
    img = syn.synth_image(rng)
    rounds = [syn.encode_round_gen(syn.synth_context(rng, r), syn.synth_answers(rng, n)) for r, n in ((1, 7), (3, 12), (10, 9))]
    rounds.append(syn.encode_round_gen(syn.synth_context(rng, 2), [[], [5000], []]))             # empty answers: no own rows at all
    slots = [0] * len(rounds)
    full = pack_units(units_from_rounds(rounds, slots), img[0][None], img[1][None], img[2][None])
    lean = pack_units(units_from_rounds(rounds, slots), img[0][None], img[1][None], img[2][None], scores_only=True, share_first_mask=False)
    pb = pack_units(units_from_rounds(rounds, slots), img[0][None], img[1][None], img[2][None], scores_only=True)
    _check(lean, rounds, scores_only=True)
    _check(pb, rounds, scores_only=True, shared_b0=True)
    assert lean.n_text_rows == full.n_text_rows - 2 * pb.n_cands and lean.c_struct().no_cls_rows == 1 and full.c_struct().no_cls_rows == 0
    assert pb.n_text_rows == lean.n_text_rows - pb.n_cands + len(rounds) and pb.n_b0_shared == len(rounds)
    for other in (lean, pb):
        assert (other.lm_labels == full.lm_labels).all() and (other.cand_lm_off == full.cand_lm_off).all()
    # the image rows never look at candidate rows at all (co-attention interval = the context rows)
    for j in pb.jobs_i2t.numpy():
        assert j[2] + j[3] <= pb.n_shared_rows
    # a unit whose masked copies differ in their first position keeps one B_0 per candidate
    odd = syn.encode_round_gen(syn.synth_context(rng, 2), syn.synth_answers(rng, 4))
    odd.positions[1, int(odd.desc[1, 2])] += 1
    pb2 = pack_units(units_from_rounds([odd], [0]), img[0][None], img[1][None], img[2][None], scores_only=True)
    assert pb2.n_b0_shared == 0
    _check(pb2, [odd], scores_only=True)


def test_packing_the_reference_made_inputs():
    g, batch = load_golden("gen100_default")
    desc = descriptors_from_masks(batch["txt_attention_mask"], batch["co_attention_mask"])
    units = units_from_flat(batch["tokens"], batch["segments"], batch["positions"], batch["mask"], desc, np.zeros(100, np.int64))
    pb = pack_units(units, g["image_feat"][None], g["image_loc"][None], g["image_mask"][None])
    r = syn.Round(batch["tokens"].numpy(), batch["segments"].numpy(), batch["positions"].numpy(), batch["mask"].numpy(), desc.numpy())
    _check(pb, [r])
    assert pb.n_text_rows == 238 + int((1 + 2 * desc[:, 3]).sum())


def test_packer_rejects_units_that_do_not_share_a_context():
    rng = np.random.RandomState(1)
    img = syn.synth_image(rng)
    r = syn.encode_round_gen(syn.synth_context(rng, 2), syn.synth_answers(rng, 4))
    r.tokens[2, 5] += 1
    try:
        pack_units(units_from_rounds([r]), img[0][None], img[1][None], img[2][None])
    except ValueError:
        return
    raise AssertionError("differing contexts must be rejected")


def test_random_batches_in_every_layout():
    """Seeded random batches (1-5 units, 1-25 candidates, answers of 0-7 tokens, rounds 1-10) through the three layouts: every
    packed row reproduces its dense position and its dense-mask key set; the layouts agree on labels and candidate offsets."""
    rng = np.random.RandomState(1234)
    img = syn.synth_image(rng)
    for trial in range(25):
        rounds = []
        for _ in range(int(rng.randint(1, 6))):
            n = int(rng.randint(1, 26))
            answers = [list(rng.randint(1000, 30522, size=int(rng.randint(0, 8)))) for _ in range(n)]
            rounds.append(syn.encode_round_gen(syn.synth_context(rng, int(rng.randint(1, 11))), answers))
        slots = [0] * len(rounds)
        args = (units_from_rounds(rounds, slots), img[0][None], img[1][None], img[2][None])
        full = pack_units(*args)
        lean = pack_units(*args, scores_only=True, share_first_mask=False)
        b0 = pack_units(*args, scores_only=True)
        # units given as row ranges of ONE stacked array per field (the reference loader's layout) take the grouped gather path
        tok, seg, pos, lab, dsc, index = syn.stack_rounds(rounds)
        stacked = units_from_flat(tok, seg, pos, lab, dsc, index)
        for u in stacked:
            u.image_slot = 0
        assert len(stacked) == len(rounds) and all(u.tokens.base is stacked[0].tokens.base for u in stacked)
        for got, kw in ((full, {}), (b0, dict(scores_only=True))):
            again = pack_units(stacked, *args[1:], **kw)
            for k in PackedBatch.INT_FIELDS + PackedBatch.FLOAT_FIELDS:
                assert torch.equal(getattr(got, k), getattr(again, k)), (trial, kw, k)
        # the vectorised packer and the per-unit loop build the same batch, tensor for tensor
        for got, kw in ((full, {}), (lean, dict(scores_only=True, share_first_mask=False)), (b0, dict(scores_only=True))):
            want = pack_units_loop(*args, **kw)
            for k in PackedBatch.INT_FIELDS + PackedBatch.FLOAT_FIELDS:
                assert torch.equal(getattr(got, k), getattr(want, k)), (trial, kw, k)
            for k in ("n_text_rows", "n_shared_rows", "cand_halo", "win_cap", "kv_cap_text", "max_q_text_self", "n_jobs_text_ctx",
                      "pairs_text_self", "pairs_i2t", "n_b0_shared"):
                assert getattr(got, k) == getattr(want, k), (trial, kw, k)
        _check(full, rounds)
        _check(lean, rounds, scores_only=True)
        shared = [len(r.desc) > 1 for r in rounds]          # a unit of one candidate keeps its own B_0 row
        _check(b0, rounds, scores_only=True, shared_b0=shared)
        assert b0.n_b0_shared == sum(shared)
        assert lean.n_text_rows == full.n_text_rows - 2 * full.n_cands
        assert b0.n_text_rows == lean.n_text_rows - sum(len(r.desc) - 1 for r, sh in zip(rounds, shared) if sh)
        for other in (lean, b0):
            assert (other.lm_labels == full.lm_labels).all() and (other.cand_lm_off == full.cand_lm_off).all()
        # the distinct-row lists reconstruct lm_rows
        assert (b0.lm_urows[b0.lm_uidx.long()] == b0.lm_rows).all() and len(torch.unique(b0.lm_urows)) == len(b0.lm_urows)
        assert int(b0.row_iv[:, 1].max()) <= b0.n_text_rows and int(b0.lm_rows.max()) < b0.n_text_rows


def test_units_backed_by_odd_arrays_pack_the_same():
    """Units whose arrays are not row ranges of a contiguous array (column-strided views, different bases per field) go through
    the stand-alone path and give the same batch."""
    rng = np.random.RandomState(9)
    img = syn.synth_image(rng)
    rounds = [syn.encode_round_gen(syn.synth_context(rng, r), syn.synth_answers(rng, n)) for r, n in ((2, 5), (6, 8))]
    want = pack_units(units_from_rounds(rounds, [0, 0]), img[0][None], img[1][None], img[2][None], scores_only=True)
    odd = []
    for r in rounds:
        wide = np.zeros((len(r.tokens), 512), np.int64)
        wide[:, ::2] = r.tokens                                    # every second column: a non-contiguous view
        odd.append(syn.Round(wide[:, ::2], np.asfortranarray(r.segments), r.positions.copy(), r.labels, r.desc))
    got = pack_units(units_from_rounds(odd, [0, 0]), img[0][None], img[1][None], img[2][None], scores_only=True)
    for k in PackedBatch.INT_FIELDS + PackedBatch.FLOAT_FIELDS:
        assert torch.equal(getattr(got, k), getattr(want, k)), k
