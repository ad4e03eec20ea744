"""GPU: ranking metrics kernel against the oracle restatement of the reference's visdial_metrics classes."""
import numpy as np
import pytest
import torch

from conftest import load_golden

pytestmark = pytest.mark.gpu
if not torch.cuda.is_available():
    pytest.skip("needs a CUDA device", allow_module_level=True)

from oracle import visdial_metrics as om  # noqa: E402
from unimm_b200.metrics import rank_metrics  # noqa: E402


def test_rank_metrics_match_oracle_on_random_scores():
    g = torch.Generator().manual_seed(0)
    scores = torch.randn(6, 10, 100, generator=g)
    gt = torch.randint(0, 100, (6, 10), generator=g)
    rel_rows = torch.tensor(np.random.RandomState(1).choice([0, 0, 0, 0.2, 0.5, 1.0], size=(6, 100)).astype(np.float32))
    rel_rows[:, 0] = 1.0
    out = rank_metrics(scores.cuda(), gt.cuda())
    want = om.sparse_metrics(scores, gt)
    for k, v in want.items():
        assert out[k] == pytest.approx(v, abs=1e-6), k
    assert out["ties"] == 0
    assert torch.equal(out["ranks"].cpu().long(), om.scores_to_ranks(scores))
    round_scores = scores[:, 3, :]
    nd = rank_metrics(round_scores.cuda(), relevance=rel_rows.cuda(), return_ranks=False)
    assert nd["ndcg"] == pytest.approx(om.ndcg(round_scores, rel_rows), abs=1e-6)


def test_rank_metrics_on_golden_config1():
    g, _ = load_golden("gen100_default")
    score = torch.from_numpy(g["seq_score"]).view(1, 100).cuda()
    out = rank_metrics(score, torch.zeros(1, dtype=torch.long).cuda(), torch.from_numpy(g["relevance"]).cuda())
    assert np.array_equal(out["ranks"].view(100).cpu().numpy(), g["ranks"])
    for k, v in zip(g["metric_names"], g["metric_values"]):
        assert out[str(k)] == pytest.approx(v, abs=1e-6), k


def test_ties_are_reported_and_ranked_stably():
    s = torch.tensor([[0.5, 0.7, 0.5, 0.1]]).cuda()
    out = rank_metrics(s, torch.tensor([2]).cuda())
    assert out["ties"] == 1 and out["ranks"].view(-1).tolist() == [2, 1, 3, 4]


def test_neural_ndcg_and_ensemble_match_reference_values():
    """Dense-annotation objective pieces (utils/rank_loss.py:518-581, val.py:152-161) against golden values from the reference."""
    import os
    from unimm_b200.rank_loss import ensemble_normalise, neural_ndcg_loss
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "rankloss.npz"))
    for p, y, want in zip(z["y_pred"], z["y_true"], z["loss"]):
        got = neural_ndcg_loss(torch.from_numpy(p).cuda(), torch.from_numpy(y).cuda()).item()
        print(f"neuralNDCG: {got:.7f} vs reference {float(want):.7f}")
        assert abs(got - float(want)) < 2e-5
    ens = ensemble_normalise(torch.from_numpy(z["ens_probs"]).cuda()).cpu().numpy()
    np.testing.assert_allclose(ens, z["ens_out"], atol=1e-6, rtol=0)


def test_neural_ndcg_gradient_matches_reference_autograd():
    """unimm_neural_ndcg_backward against ``y_pred.grad`` after ``neuralNDCG_transposed(y_pred, y_true).backward()`` of the unmodified
    reference (tests/golden/rankloss_grad.npz), and the softmax(NSP)[:, 0] chain onto the logits against autograd."""
    import os
    import ctypes as C
    from unimm_b200._lib import check, lib, ptr
    from unimm_b200.rank_loss import neural_ndcg_loss_backward
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "rankloss_grad.npz"))
    for k, (p, y, want, loss) in enumerate(zip(z["y_pred"], z["y_true"], z["grad"], z["loss"])):
        d, ndcg = neural_ndcg_loss_backward(torch.from_numpy(p).cuda(), torch.from_numpy(y).cuda())
        err = np.abs(d.cpu().numpy() - want).max() / np.abs(want).max()
        valid = (np.power(2.0, y) - 1).sum(-1) != 0
        got_loss = -float(ndcg.cpu().numpy()[valid].sum() / valid.sum())
        print(f"neuralNDCG gradient case {k}: max |err| / max |grad| = {err:.2e}; loss {got_loss:.7f} vs {float(loss):.7f}")
        assert err < 2e-4 and abs(got_loss - float(loss)) < 2e-5
        assert np.abs(d.cpu().numpy()[~valid]).max(initial=0.0) == 0.0
    # y_pred = softmax(nsp)[:, 0] and its backward (accumulating)
    g = torch.Generator().manual_seed(1)
    logits = torch.randn(100, 2, generator=g)
    dp = torch.randn(100, generator=g)
    x = logits.clone().requires_grad_()
    torch.softmax(x, -1)[:, 0].backward(dp)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    dl, dlog, p0 = logits.cuda(), torch.ones(100, 2, device="cuda"), torch.empty(100, device="cuda")
    check(lib.unimm_t_nsp_prob0(ptr(dl), 100, ptr(p0), st))
    check(lib.unimm_t_nsp_prob0_backward(ptr(dl), ptr(dp.cuda()), 100, ptr(dlog), st))
    assert (p0.cpu() - torch.softmax(logits, -1)[:, 0]).abs().max().item() < 1e-6
    assert (dlog.cpu() - 1.0 - x.grad).abs().max().item() < 1e-6
