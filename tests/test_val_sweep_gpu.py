"""GPU: the generative ranking sweep end to end (pack -> prefix-shared forward through the host-buffer C ABI -> GPU ranks /
metrics -> EvalAI records) against the oracle's metric functions on the same scores."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

if not torch.cuda.is_available():
    pytest.skip("needs a CUDA device", allow_module_level=True)

from oracle import visdial_metrics as om  # noqa: E402
from unimm_b200.engine import Engine  # noqa: E402
from unimm_b200.val_sweep import gpu_metrics, packed_scorer, run_sweep, synthetic_items  # noqa: E402
from unimm_b200.weights import random_state_dict  # noqa: E402


def test_sweep_metrics_and_ranks_match_the_oracle(full_cfg):
    items = synthetic_items(range(3))
    eng = Engine(full_cfg, random_state_dict(full_cfg, 0), precision="fp16", max_sequences=3 * 52)
    try:
        dev = torch.device("cuda", 0)
        res = run_sweep(items, packed_scorer(eng), 0, 1, images_per_step=2, metrics_fn=lambda s, g, ns, r: gpu_metrics(s, g, ns, r, dev))
    finally:
        eng.close()
    s = res["scores"]
    assert s.shape == (3, 10, 100) and torch.isfinite(s).all()
    gt = torch.zeros(3, 10, dtype=torch.long)
    want = om.sparse_metrics(s, gt)
    for k, v in want.items():
        assert abs(res["metrics"][k] - v) < 1e-6, k        # the oracle averages in fp32
    ann_scores = torch.stack([s[i, it.relevance_round] for i, it in enumerate(items)])
    rel = torch.from_numpy(np.stack([it.relevance for it in items]))
    assert abs(res["metrics"]["ndcg"] - om.ndcg(ann_scores, rel)) < 1e-6
    ranks = om.scores_to_ranks(s)
    assert res["metrics"]["ties"] == 0
    for rec in res["predictions"]:
        i, j = rec["image_id"], rec["round_id"] - 1
        assert rec["ranks"] == ranks[i, j].tolist()
    print("sweep metrics:", {k: round(float(v), 4) for k, v in res["metrics"].items()})


def test_nsp_sweep_matches_reference_ranking_rule(full_cfg):
    """val.py's discriminative ranking (NSP probability, per-round min-max normalisation, ensemble sum) through run_sweep: the scores
    equal the rule applied by hand to the engine's own NSP logits, and a two-member 'ensemble' of the same model changes nothing
    but the scale."""
    import numpy as np

    from conftest import golden_state_dict
    from unimm_b200.engine import Engine
    from unimm_b200.val_sweep import NspScorer, gpu_metrics, run_sweep, synthetic_items
    items = synthetic_items(range(2), n_candidates=12, n_rounds=3, mode="dis")
    eng = Engine(full_cfg, golden_state_dict(full_cfg, 0, False), precision="fp16", max_sequences=64)
    dev = eng.device
    res = run_sweep(items, NspScorer([eng]), images_per_step=1, metrics_fn=lambda s, g, ns, r: gpu_metrics(s, g, ns, r, dev))
    res2 = run_sweep(items, NspScorer([eng, eng]), images_per_step=2, metrics_fn=lambda s, g, ns, r: gpu_metrics(s, g, ns, r, dev))
    assert res["scores"].shape == (2, 3, 12)
    np.testing.assert_allclose(res2["scores"].numpy(), 2 * res["scores"].numpy(), rtol=1e-5, atol=1e-7)
    assert torch.equal(res["ranks"].cpu(), res2["ranks"].cpu())
    # by hand for image 0, round 1
    it, r = items[0], items[0].rounds[1]
    o = eng.forward(torch.from_numpy(r.tokens), torch.from_numpy(r.segments), torch.from_numpy(r.positions), torch.from_numpy(r.desc),
                    torch.from_numpy(it.feat)[None], torch.from_numpy(it.loc)[None], torch.from_numpy(it.mask)[None],
                    feat_index=torch.zeros(12, dtype=torch.int32), want=("nsp_scores",))
    p = torch.softmax(o["nsp_scores"], 1)[:, 0].cpu()
    e = (p - p.min()) / (p.max() - p.min())
    np.testing.assert_allclose(res["scores"][0, 1].numpy(), (e / e.sum()).numpy(), rtol=1e-4, atol=1e-6)
    eng.close()
