import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "slow: CPU test that takes more than ~20 s")


def load_golden(name):
    """Load a fixture written by tests/golden/make_golden.py back into the flattened batch layout."""
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    g = {k: z[k] for k in z.files}
    S = g["tokens"].shape[1]
    att = np.unpackbits(g["txt_attention_mask"], axis=-1)[..., :S]
    att_t = torch.from_numpy(att)
    att_t = att_t.long() if bool(g["txt_attention_mask_is_long"]) else att_t.bool()
    n = g["tokens"].shape[0]
    R = g["image_feat"].shape[0]
    batch = {
        "tokens": torch.from_numpy(g["tokens"]), "segments": torch.from_numpy(g["segments"]),
        "positions": torch.from_numpy(g["positions"]), "sep_indices": torch.from_numpy(g["sep_indices"]),
        "mask": torch.from_numpy(g["mask"]), "weights": torch.from_numpy(g["weights"]),
        "txt_attention_mask": att_t,
        "co_attention_mask": torch.from_numpy(g["co_txt_mask"]).long().unsqueeze(1).repeat(1, R, 1),
        "image_feat": torch.from_numpy(g["image_feat"]).unsqueeze(0).expand(n, -1, -1).contiguous(),
        "image_loc": torch.from_numpy(g["image_loc"]).unsqueeze(0).expand(n, -1, -1).contiguous(),
        "image_mask": torch.from_numpy(g["image_mask"]).unsqueeze(0).expand(n, -1).contiguous(),
    }
    return g, batch


_SD_CACHE = {}


def golden_state_dict(cfg, seed, perturbed):
    from unimm_b200.weights import random_state_dict
    key = (id(type(cfg)), cfg.num_hidden_layers, cfg.v_num_hidden_layers, int(seed), bool(perturbed))
    if key not in _SD_CACHE:
        _SD_CACHE.clear()                      # 1 GB each: keep one
        _SD_CACHE[key] = random_state_dict(cfg, int(seed), bool(perturbed))
    return _SD_CACHE[key]


@pytest.fixture(scope="session")
def full_cfg():
    from unimm_b200.config import DEFAULT_CONFIG_PATH, ViLBertConfig
    return ViLBertConfig.from_json_file(DEFAULT_CONFIG_PATH)


def load_golden_multi_image(name):
    """Fixtures whose sequences belong to several images (train240_*): the feature / target blocks are regenerated from the stored
    seeds (oracle.encode_inputs.synth_image, a seeded Dirichlet) with the stored masking decisions of the reference's
    encode_image_input applied; returns (g, batch, per-image blocks)."""
    from oracle import encode_inputs as enc
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    g = {k: z[k] for k in z.files}
    S = g["tokens"].shape[1]
    att = torch.from_numpy(np.unpackbits(g["txt_attention_mask"], axis=-1)[..., :S])
    att = att.long() if bool(g["txt_attention_mask_is_long"]) else att.bool()
    feats, targets = [], []
    for i, seed in enumerate(g["image_seeds"]):
        f, _, _ = enc.synth_image(np.random.RandomState(int(seed)))
        f = f.clone()
        f[torch.from_numpy(g["image_zeroed"][i])] = 0
        feats.append(f)
        targets.append(torch.from_numpy(np.random.RandomState(int(seed) + 1000).dirichlet(np.ones(1601), size=37).astype(np.float32)))
    R = feats[0].shape[0]
    blocks = {"image_feat": torch.stack(feats), "image_loc": torch.from_numpy(g["image_loc"]), "image_mask": torch.from_numpy(g["image_mask"]),
              "image_target": torch.stack(targets), "image_label": torch.from_numpy(g["image_label"])}
    batch = {"tokens": torch.from_numpy(g["tokens"]), "segments": torch.from_numpy(g["segments"]), "positions": torch.from_numpy(g["positions"]),
             "sep_indices": torch.from_numpy(g["sep_indices"]), "mask": torch.from_numpy(g["mask"]), "weights": torch.from_numpy(g["weights"]),
             "txt_attention_mask": att, "co_attention_mask": torch.from_numpy(g["co_txt_mask"]).long().unsqueeze(1).repeat(1, R, 1),
             "seq_image": torch.from_numpy(g["seq_image"])}
    return g, batch, blocks
