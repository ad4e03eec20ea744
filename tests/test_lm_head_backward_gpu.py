"""GPU: backward of the fused LM head + L / UL loss (unimm_k_lm_head_backward) against torch.autograd in fp64 on the SAME 16-bit
operands (the rounding of h and E to the operand format is the forward's, not the backward's)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

if not torch.cuda.is_available():
    pytest.skip("needs a CUDA device", allow_module_level=True)

from unimm_b200.lm_head_grad import lm_head_backward  # noqa: E402


def reference(h, E, b, labels, w, scale):
    """The reference's loss (models/vilbert_dialog.py:1577-1595) in fp64 with autograd."""
    h = h.double().requires_grad_(True)
    E = E.double().requires_grad_(True)
    b = b.double().requires_grad_(True)
    z = h @ E.t() + b
    logp_all = torch.log_softmax(z, -1)
    lp = logp_all.gather(1, labels.view(-1, 1).long())[:, 0]
    l_rows, ul_rows = w > 0, w == -1
    loss = -(w[l_rows].double() * lp[l_rows]).sum() - torch.log(torch.clamp(1.0 - lp[ul_rows].exp(), min=1e-6)).sum()
    (loss * scale).backward()
    return h.grad, E.grad, b.grad, lp.detach()


@pytest.mark.parametrize("precision,tol", [("fp16", 4e-3), ("bf16", 3e-2)])
@pytest.mark.parametrize("n", [300, 64])
def test_lm_head_backward_matches_autograd(precision, tol, n):
    g = torch.Generator().manual_seed(5 + n)
    V, K = 30522, 768
    dt = torch.float16 if precision == "fp16" else torch.bfloat16
    h = torch.randn(n, K, generator=g).to(dt).float()                   # values exactly representable in the operand format
    E = 0.02 * torch.randn(V, K, generator=g)
    E[:50] *= 30.0                                                      # a few confident tokens: p(label) far from 1 / V, UL rows with real weight
    E = E.to(dt).float()
    b = 0.1 * torch.randn(V, generator=g)
    labels = torch.randint(0, V, (n,), generator=g)
    labels[: n // 3] = torch.randint(0, 50, (n // 3,), generator=g)
    w = torch.ones(n)
    w[1::3] = -1.0                                                      # unlikelihood rows
    w[2::7] = 0.0                                                       # rows that carry no loss at all
    w[5::11] = 2.0
    scale = 1.0 / float((w != 0).sum())
    out = lm_head_backward(h.cuda(), E.cuda(), b, labels, w, grad_scale=scale, precision=precision)
    dH, dE, db, lp = reference(h, E, b, labels, w, scale)
    for name, mine, ref in (("dH", out["dH"], dH), ("dE", out["dE"], dE), ("dbias", out["dbias"], db)):
        err = (mine.cpu().double() - ref).abs().max().item() / ref.abs().max().item()
        print(f"[{precision}] n={n} {name}: max |err| / max |ref| = {err:.3e}")
        assert err < tol, name
    assert (out["logp"].cpu().double() - lp).abs().max().item() < 2e-3
    # rows without a loss get no gradient
    assert out["dH"][(w == 0).nonzero().view(-1).cuda()].abs().max().item() == 0.0


@pytest.mark.parametrize("precision,tol", [("fp16", 3e-3), ("bf16", 2e-2)])
@pytest.mark.parametrize("shape", [(300, 3072, 768), (1000, 768, 3072), (77, 2304, 768)])
def test_linear_backward_matches_autograd(precision, tol, shape):
    """dgrad / wgrad / db of one projection (unimm_k_linear_backward) against fp64 autograd; the incoming gradient is tiny (1e-6:
    below fp16's normal range) to exercise the power-of-two scaling."""
    from unimm_b200.lm_head_grad import linear_backward
    M, N, K = shape
    g = torch.Generator().manual_seed(M + N)
    dt = torch.float16 if precision == "fp16" else torch.bfloat16
    X = torch.randn(M, K, generator=g).to(dt).float()
    W = (0.02 * torch.randn(N, K, generator=g)).to(dt).float()
    dY = 1e-6 * torch.randn(M, N, generator=g)
    out = linear_backward(dY.cuda(), X, W, precision=precision)
    dX = dY.double() @ W.double()
    dW = dY.double().t() @ X.double()
    db = dY.double().sum(0)
    for name, mine, ref in (("dX", out["dX"], dX), ("dW", out["dW"], dW), ("db", out["db"], db)):
        err = (mine.cpu().double() - ref).abs().max().item() / ref.abs().max().item()
        print(f"[{precision}] {shape} {name}: max |err| / max |ref| = {err:.3e}")
        assert err < (1e-5 if name == "db" else tol), name


@pytest.mark.parametrize("precision,tol", [("fp16", 1e-2), ("bf16", 6e-2)])
def test_whole_lm_head_backward_matches_autograd(precision, tol):
    """cls.predictions (transform dense -> erf-GELU -> LayerNorm -> tied decoder + bias) under the L / UL loss: the chained device
    kernels against fp64 autograd of the same module — gradient w.r.t. the head's input rows and every parameter of the head."""
    from unimm_b200.lm_head_grad import lm_head_block_backward
    g = torch.Generator().manual_seed(21)
    n, V, K = 200, 30522, 768
    dt = torch.float16 if precision == "fp16" else torch.bfloat16
    x = torch.randn(n, K, generator=g).to(dt).float()
    P = {"transform.dense.weight": (0.03 * torch.randn(K, K, generator=g)).to(dt).float(), "transform.dense.bias": 0.05 * torch.randn(K, generator=g),
         "transform.LayerNorm.weight": 1.0 + 0.1 * torch.randn(K, generator=g), "transform.LayerNorm.bias": 0.1 * torch.randn(K, generator=g),
         "decoder.weight": 0.02 * torch.randn(V, K, generator=g), "bias": 0.1 * torch.randn(V, generator=g)}
    P["decoder.weight"][:40] *= 25.0
    P["decoder.weight"] = P["decoder.weight"].to(dt).float()
    labels = torch.randint(0, V, (n,), generator=g)
    labels[: n // 3] = torch.randint(0, 40, (n // 3,), generator=g)
    w = torch.ones(n)
    w[1::3] = -1.0
    scale = 1.0 / float((w != 0).sum())
    out = lm_head_block_backward(x.cuda(), P, labels, w, grad_scale=scale, precision=precision)
    # fp64 reference of the same module
    xr = x.double().requires_grad_(True)
    R = {k: v.double().requires_grad_(True) for k, v in P.items()}
    t = xr @ R["transform.dense.weight"].t() + R["transform.dense.bias"]
    gl = t * 0.5 * (1.0 + torch.erf(t / 2.0 ** 0.5))
    mu, var = gl.mean(-1, keepdim=True), gl.var(-1, unbiased=False, keepdim=True)
    h = (gl - mu) / torch.sqrt(var + 1e-12) * R["transform.LayerNorm.weight"] + R["transform.LayerNorm.bias"]
    z = h @ R["decoder.weight"].t() + R["bias"]
    lp = torch.log_softmax(z, -1).gather(1, labels.view(-1, 1))[:, 0]
    loss = -(w[w > 0].double() * lp[w > 0]).sum() - torch.log(torch.clamp(1.0 - lp[w == -1].exp(), min=1e-6)).sum()
    (loss * scale).backward()
    refs = {"dx": xr.grad, **{k: v.grad for k, v in R.items()}}
    for name, ref in refs.items():
        err = (out[name].cpu().double() - ref).abs().max().item() / ref.abs().max().item()
        print(f"[{precision}] whole LM head, {name}: max |err| / max |ref| = {err:.3e}")
        assert err < tol, name
