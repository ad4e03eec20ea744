"""CPU: host-side logic — checkpoint layout, descriptors <-> dense masks, synthetic encoders, C-ABI exports."""
import os
import re

import numpy as np
import torch

from conftest import ROOT, load_golden
from oracle import encode_inputs as enc
from unimm_b200 import synthetic as syn
from unimm_b200.config import DEFAULT_CONFIG_PATH, ViLBertConfig, tiny_config
from unimm_b200.descriptors import dense_co_mask, dense_text_mask, descriptors_from_masks
from unimm_b200.weights import param_shapes, random_state_dict


def test_checkpoint_layout_is_the_references(full_cfg):
    shapes = param_shapes(full_cfg)
    assert len(shapes) == 535
    distinct = sum(int(np.prod(s)) for k, s in shapes.items() if k != "cls.predictions.decoder.weight")
    assert distinct == 250_090_109                                # parameters counted once (tied decoder)
    assert sum(int(np.prod(s)) for s in shapes.values()) == 273_531_005
    assert list(shapes)[0] == "bert.embeddings.word_embeddings.weight"
    assert "bert.encoder.c_layer.5.biOutput.q_dense2.bias" in shapes       # unused-but-present keys are kept


def test_random_state_dict_is_deterministic():
    cfg = tiny_config()
    a, b = random_state_dict(cfg, 5, True), random_state_dict(cfg, 5, True)
    assert all(torch.equal(a[k], b[k]) for k in a)
    assert a["cls.predictions.decoder.weight"] is a["bert.embeddings.word_embeddings.weight"]
    d = random_state_dict(cfg, 5, False)
    assert float(d["bert.encoder.layer.0.output.dense.bias"].abs().max()) == 0.0
    assert float(a["bert.encoder.layer.0.output.dense.bias"].abs().max()) > 0.0


def test_layer_schedule_matches_reference_order(full_cfg):
    s = full_cfg.layer_schedule()
    assert s[:7] == [("t", i) for i in range(6)] + [("c", 0)]
    assert s[7:10] == [("v", 0), ("t", 6), ("c", 1)]
    assert s[-2:] == [("v", 5), ("t", 11)] and len(s) == 24


def test_descriptors_roundtrip_reference_masks():
    for name in ("gen8_default", "dis8_perturbed", "train6_perturbed"):
        g, batch = load_golden(name)
        desc = descriptors_from_masks(batch["txt_attention_mask"], batch["co_attention_mask"], verify=True)
        assert torch.equal(dense_text_mask(desc, 256), batch["txt_attention_mask"].bool())
        assert torch.equal(dense_co_mask(desc, 256), batch["co_attention_mask"][:, 0, :])
        labelled = (batch["mask"] != -1)
        gen = desc[:, 0] == 0
        # generative rows: every position of the masked copy [L, T) is labelled
        for b in torch.nonzero(gen)[:, 0].tolist():
            _, ctx, L, last = desc[b].tolist()
            assert labelled[b, L:L + last].all() and ctx == L - last


def test_descriptor_derivation_rejects_other_masks():
    g, batch = load_golden("gen8_default")
    bad = batch["txt_attention_mask"].clone()
    bad[0, 5, 250] = True        # a context row may not see the answer
    try:
        descriptors_from_masks(bad, batch["co_attention_mask"])
    except NotImplementedError:
        return
    raise AssertionError("a non-reference mask pattern must be rejected")


def test_synthetic_round_equals_oracle_encoder():
    rng = np.random.RandomState(3)
    context, answers = syn.synth_context(rng, round_id=4), syn.synth_answers(rng, 6)
    r = syn.encode_round_gen(context, answers)
    d = syn.encode_round_dis(context, answers)
    for j, ans in enumerate(answers):
        tok, seg, pos, _, lab, _, att, co = enc.encode_gen(context + [ans], 1, mask_prob=0, rng=np.random.RandomState(0))
        assert np.array_equal(r.tokens[j], tok[0].numpy()) and np.array_equal(r.segments[j], seg[0].numpy())
        assert np.array_equal(r.positions[j], pos[0].numpy()) and np.array_equal(r.labels[j], lab[0].numpy())
        desc = torch.from_numpy(r.desc[j:j + 1])
        assert torch.equal(dense_text_mask(desc, 256)[0], att[0].bool()) and torch.equal(dense_co_mask(desc, 256)[0], co[0])
        tok, seg, pos, _, lab, _, att, co = enc.encode_dis(context + [ans], 1, mask_prob=0, rng=np.random.RandomState(0))
        assert np.array_equal(d.tokens[j], tok[0].numpy()) and np.array_equal(d.segments[j], seg[0].numpy())
        assert np.array_equal(d.positions[j], pos[0].numpy()) and np.array_equal(d.labels[j], lab[0].numpy())
        desc = torch.from_numpy(d.desc[j:j + 1])
        assert torch.equal(dense_text_mask(desc, 256)[0], att[0].bool()) and torch.equal(dense_co_mask(desc, 256)[0], co[0])


def test_config1_context_has_239_positions():
    rng = np.random.RandomState(0)
    r = syn.encode_round_gen(syn.synth_context(rng, 10), syn.synth_answers(rng, 100))
    assert (r.desc[:, 1] == 239).all() and r.desc[:, 3].min() >= 2 and r.desc[:, 3].max() <= 8
    assert ((r.desc[:, 2] + r.desc[:, 3]) <= 255).all()


def test_vectorised_round_encoder_equals_the_per_candidate_statement():
    rng = np.random.RandomState(3)
    for r, n in ((1, 7), (10, 100), (5, 33), (2, 1)):
        c, a = syn.synth_context(rng, r), syn.synth_answers(rng, n)
        if n == 7:
            a[2] = []                                        # an empty answer: only [SEP] / one [MASK]
        x, y = syn.encode_round_gen(c, a), syn._encode_round_gen_loop(c, a)
        for f in ("tokens", "segments", "positions", "labels", "desc"):
            assert np.array_equal(getattr(x, f), getattr(y, f)) and getattr(x, f).dtype == getattr(y, f).dtype, (f, r, n)


def test_library_loads_and_exports_every_declared_symbol():
    """No compute without a GPU: only that the in-tree .so loads and matches include/unimm_b200.h."""
    from unimm_b200 import _lib
    header = open(os.path.join(ROOT, "include", "unimm_b200.h")).read()
    declared = set(re.findall(r"\b(unimm_[a-z0-9_]+)\s*\(", header))
    assert declared, "header parse failed"
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    for name in declared:
        assert hasattr(_lib.lib, name), name
    assert _lib.lib.unimm_abi_version() == 3
    assert _lib.LIB_PATH.startswith(ROOT)          # in-tree, so the driver sees it loaded


def test_drop_in_module_has_reference_state_dict_keys(full_cfg):
    from unimm_b200.visual_dialog_encoder import VisualDialogEncoder
    m = VisualDialogEncoder(DEFAULT_CONFIG_PATH)
    keys = list(m.state_dict().keys())
    assert keys == ["bert_pretrained." + k for k in param_shapes(full_cfg)]
    sd = m.state_dict()
    assert sd["bert_pretrained.cls.predictions.decoder.weight"].data_ptr() == \
        sd["bert_pretrained.bert.embeddings.word_embeddings.weight"].data_ptr()


def test_product_never_imports_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "unimm_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f


def test_reference_arm_prints_one_json_line():
    """bench.py --impl reference (the driver's reference arm) on a 4-candidate sample: exactly one line on stdout, carrying the
    contract's keys; `kind` says whether the unmodified reference module (baseline/_ref or /root/reference present) or the oracle
    port was timed."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    p = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--ref-candidates", "4"], capture_output=True, text=True, timeout=600, cwd=root)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, p.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "candidates_scored_per_sec" and d["value"] > 0 and d["higher_is_better"] is True
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
    from ref_callers import reference_root
    assert d["cpu_baseline"]["kind"] == ("reference" if reference_root() is not None else "port")
