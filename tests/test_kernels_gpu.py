"""GPU: each CUDA kernel, called through the C ABI, against a plain PyTorch fp32 statement of the same op."""
import ctypes as C
import math

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

if not torch.cuda.is_available():
    pytest.skip("needs a CUDA device", allow_module_level=True)

from unimm_b200 import _lib  # noqa: E402
from unimm_b200._lib import check, lib, ptr  # noqa: E402
from unimm_b200.descriptors import dense_co_mask, dense_text_mask  # noqa: E402

DEV = torch.device("cuda", 0)


def stream():
    return C.c_void_p(torch.cuda.current_stream(DEV).cuda_stream)


def gelu(x):
    return x * 0.5 * (1.0 + torch.erf(x / math.sqrt(2.0)))


def rnd(*shape, scale=1.0, seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(DEV)


# ----------------------------------------------------------------------------------------------- fp32 GEMM
@pytest.mark.parametrize("M,N,K,act,res", [(256, 768, 768, 0, True), (300, 3072, 768, 1, False), (37 * 3, 1024, 2048, 0, True),
                                           (130, 1601, 1024, 0, False), (64, 30522, 768, 0, False)])
def test_gemm_f32(M, N, K, act, res):
    A, W, b = rnd(M, K, seed=1), rnd(N, K, scale=0.05, seed=2), rnd(N, seed=3)
    R = rnd(M, N, seed=4) if res else None
    out = torch.empty(M, N, device=DEV)
    check(lib.unimm_k_gemm_f32(ptr(A), K, ptr(W), K, M, N, K, ptr(b), ptr(R), N, act, ptr(out), N, stream()))
    ref = A.double() @ W.double().t() + b.double()
    if act == 1:
        ref = gelu(ref)
    if res:
        ref = ref + R.double()
    err = (out.double() - ref).abs().max().item()
    print(f"gemm_f32 {M}x{N}x{K} act={act} res={res}: max abs err {err:.3e}")
    assert err < 2e-4


# ----------------------------------------------------------------------------------------------- tcgen05 GEMM
LP = {"bf16": (torch.bfloat16, 0), "fp16": (torch.float16, 1)}


@pytest.mark.parametrize("lp", ["bf16", "fp16"])
@pytest.mark.parametrize("M,N,K,act,res,tile_n", [
    (128, 256, 64, 0, False, 256),       # one tile, one k-block: descriptor / swizzle sanity
    (128, 128, 128, 0, False, 128),
    (256, 768, 768, 0, True, 256),       # out-proj shape with residual
    (512, 2304, 768, 0, False, 256),     # QKV
    (300, 3072, 768, 1, False, 256),     # FFN-1 + GELU, ragged M
    (256, 768, 3072, 0, True, 256),      # FFN-2, long K
    (37 * 5, 1024, 2048, 0, True, 128),  # image embedding shape, ragged M
    (130, 1601, 1024, 0, False, 128),    # image decoder: ragged N
    (40000, 768, 768, 0, False, 256),    # many tiles per CTA: exercises ring + accumulator phases
])
def test_gemm_umma(M, N, K, act, res, tile_n, lp):
    dt, kind = LP[lp]
    A = rnd(M, K, seed=1).to(dt)
    W = rnd(N, K, scale=0.05, seed=2).to(dt)
    b = rnd(N, seed=3)
    R = rnd(M, N, seed=4) if res else None
    o32 = torch.zeros(M, N, device=DEV)
    o16 = torch.zeros(M, N, device=DEV, dtype=dt)
    check(lib.unimm_k_gemm_lp(ptr(A), K, ptr(W), K, M, N, K, ptr(b), ptr(R), N, act, ptr(o32), N, ptr(o16), N, tile_n, 0, kind,
                              stream()))
    torch.cuda.synchronize()
    ref = A.double() @ W.double().t() + b.double()
    if act == 1:
        ref = gelu(ref)
    if res:
        ref = ref + R.double()
    err = (o32.double() - ref).abs().max().item()
    err16 = (o16.double() - ref).abs().max().item()
    print(f"gemm_umma[{lp}] {M}x{N}x{K} act={act} res={res}: fp32-out err {err:.3e}, 16-bit-out err {err16:.3e}")
    assert err < 2e-3 * max(1.0, math.sqrt(K / 768))     # operands are exactly representable: only fp32 accumulation order
    assert err16 < 0.05 * max(1.0, ref.abs().max().item() / 4)


@pytest.mark.parametrize("lp", ["bf16", "fp16"])
@pytest.mark.parametrize("M,N,K,act,tile_n", [
    (128, 256, 64, 0, 256), (300, 3072, 768, 1, 256), (300, 3072, 768, 3, 256), (512, 2304, 768, 0, 256), (37 * 5, 1024, 1024, 1, 128),
    (130, 1600, 1024, 0, 128),          # N % 32 == 0 but not a multiple of the tile: ragged last tile
    (40000, 768, 768, 0, 256),
])
def test_gemm_umma_fragment_epilogue(M, N, K, act, tile_n, lp):
    """16-bit-output GEMM with fragment-ordered weights (no shared-memory transpose); act 3 = the 1-SFU tanh-form GELU."""
    dt, kind = LP[lp]
    A = rnd(M, K, seed=1).to(dt)
    W = rnd(N, K, scale=0.05, seed=2).to(dt)
    b = rnd(N, seed=3)
    Wp = torch.empty_like(W)
    check(lib.unimm_k_permute_w(ptr(W), ptr(Wp), N, K, 1, stream()))
    o16 = torch.zeros(M, N, device=DEV, dtype=dt)
    check(lib.unimm_k_gemm_lp(ptr(A), K, ptr(Wp), K, M, N, K, ptr(b), None, 0, act, None, 0, ptr(o16), N, tile_n, 0, kind | 0x100, stream()))
    torch.cuda.synchronize()
    ref = A.double() @ W.double().t() + b.double()
    if act in (1, 3):
        ref = gelu(ref)
    err16 = (o16.double() - ref).abs().max().item()
    print(f"gemm_umma_frag[{lp}] {M}x{N}x{K} act={act}: 16-bit-out err {err16:.3e}")
    assert err16 < 0.05 * max(1.0, ref.abs().max().item() / 4)
    if lp == "fp16":   # tighter: half an fp16 ulp of the largest value plus the GELU approximation
        assert err16 < ref.abs().max().item() * (2 ** -11 + (3e-4 if act == 3 else 1e-5)) + 1e-5


def test_gemm_umma_persistent_few_ctas():
    """Force 3 CTAs over 60 tiles so every CTA wraps the smem ring and both accumulators many times."""
    M, N, K = 128 * 10, 256 * 6, 320
    A, W = rnd(M, K, seed=5).to(torch.bfloat16), rnd(N, K, scale=0.05, seed=6).to(torch.bfloat16)
    o32 = torch.zeros(M, N, device=DEV)
    check(lib.unimm_k_gemm_lp(ptr(A), K, ptr(W), K, M, N, K, None, None, 0, 0, ptr(o32), N, None, 0, 256, 3, 0, stream()))
    ref = A.double() @ W.double().t()
    assert (o32.double() - ref).abs().max().item() < 2e-3


@pytest.mark.parametrize("lp", ["bf16", "fp16"])
@pytest.mark.parametrize("res16", [False, True])
@pytest.mark.parametrize("M,N,K,inplace", [
    (128, 768, 64, False),         # one row block, one k-block, 3-CTA cluster
    (128, 1024, 128, False),       # 4-CTA cluster
    (300, 768, 768, True),         # out-proj shape, ragged M, residual stream updated in place
    (37 * 7, 1024, 1024, True),    # image stream
    (256, 768, 3072, False),       # FFN-2: long K
    (128 * 150 + 5, 768, 768, True),   # more row blocks than clusters: accumulator / statistics-slot phases wrap many times
    (128 * 101, 1024, 1024, False),
])
def test_gemm_ln_cluster(M, N, K, inplace, res16, lp):
    """LayerNorm(A W^T + b + R) * gamma + beta from the cluster-fused kernel vs float64 torch (reference :422-426).
    res16: the residual is a 16-bit tensor added on the tensor core (identity k-blocks) instead of an fp32 one."""
    dt, kind = LP[lp]
    A = rnd(M, K, seed=1).to(dt)
    W = rnd(N, K, scale=0.05, seed=2).to(dt)
    b = rnd(N, seed=3)
    R = rnd(M, N, seed=4) + 0.3          # non-zero row mean
    R[:, 5] += 8.0                       # an outlier feature column, as trained BERT residual streams have
    if res16:
        R = R.to(dt).float()             # exactly representable, so the reference sees the same residual
    gamma, beta = 1.0 + 0.1 * rnd(N, seed=5), 0.1 * rnd(N, seed=6)
    ref = torch.nn.functional.layer_norm(A.double() @ W.double().t() + b.double() + R.double(), (N,), gamma.double(), beta.double(),
                                         eps=1e-12)
    Wp = torch.empty_like(W)     # rows in the order the fused kernel's TMEM fragments want (done once at weight-load time)
    check(lib.unimm_k_permute_w_ln(ptr(W), ptr(Wp), N, K, stream()))
    o32 = R.clone() if (inplace and not res16) else torch.zeros(M, N, device=DEV)
    o16 = R.to(dt) if (inplace and res16) else torch.zeros(M, N, device=DEV, dtype=dt)
    R16 = o16 if inplace else R.to(dt)
    if res16:
        check(lib.unimm_k_gemm_ln_lp(ptr(A), K, ptr(Wp), K, M, N, K, ptr(b), None, 0, ptr(R16), N, ptr(gamma), ptr(beta),
                                     ptr(o32), N, ptr(o16), N, kind, stream()))
    else:
        check(lib.unimm_k_gemm_ln_lp(ptr(A), K, ptr(Wp), K, M, N, K, ptr(b), ptr(o32) if inplace else ptr(R), N, None, 0, ptr(gamma),
                                     ptr(beta), ptr(o32), N, ptr(o16), N, kind, stream()))
    torch.cuda.synchronize()
    err = (o32.double() - ref).abs().max().item()
    err16 = (o16.double() - ref).abs().max().item()
    print(f"gemm_ln[{lp}] {M}x{N}x{K} inplace={inplace} res16={res16}: fp32-out err {err:.3e}, 16-bit-out err {err16:.3e}")
    assert err < 2e-3 * max(1.0, math.sqrt(K / 768))
    assert err16 < 0.05


@pytest.mark.parametrize("lp", ["bf16", "fp16"])
@pytest.mark.parametrize("rows,V", [(200, 30522), (8321, 5003)])     # the second: CTA pairs with the W tile multicast, odd row-block count
def test_lm_head_lse(lp, rows, V):
    dt, kind = LP[lp]
    K = 768
    Hm = rnd(rows, K, seed=7).to(dt)
    E = rnd(V, K, scale=0.02, seed=8).to(dt)
    bias = rnd(V, scale=0.02, seed=9)
    labels = torch.randint(0, V, (rows,), generator=torch.Generator().manual_seed(3)).to(torch.int32)
    labels[0], labels[1] = 0, V - 1
    labels = labels.to(DEV)
    tiles = 2 * ((V + 255) // 256)      # two column halves per 256-wide vocabulary tile
    partials = torch.zeros(rows, tiles, 2, device=DEV)
    lab_logit = torch.zeros(rows, device=DEV)
    logp, ul = torch.zeros(rows, device=DEV), torch.zeros(rows, device=DEV)
    check(lib.unimm_k_lm_head_lp(ptr(Hm), K, ptr(E), K, rows, V, K, ptr(bias), ptr(labels), ptr(partials), ptr(lab_logit),
                                 ptr(logp), ptr(ul), kind, stream()))
    logits = Hm.double() @ E.double().t() + bias.double()
    ref = torch.log_softmax(logits, -1).gather(1, labels.long()[:, None])[:, 0]
    ref_ul = torch.log(torch.clamp(1.0 - torch.softmax(logits, -1), min=1e-6)).gather(1, labels.long()[:, None])[:, 0]
    print("lm head: logp err", (logp.double() - ref).abs().max().item(), "ul err", (ul.double() - ref_ul).abs().max().item())
    assert (logp.double() - ref).abs().max().item() < 1e-4
    assert (ul.double() - ref_ul).abs().max().item() < 1e-4


# ----------------------------------------------------------------------------------------------- LayerNorm
@pytest.mark.parametrize("H", [768, 1024])
def test_layernorm(H):
    rows = 1000
    x, g, b = rnd(rows, H, scale=3.0, seed=1) + 0.5, 1 + 0.1 * rnd(H, seed=2), 0.1 * rnd(H, seed=3)
    y32 = torch.empty(rows, H, device=DEV)
    y16 = torch.empty(rows, H, device=DEV, dtype=torch.bfloat16)
    check(lib.unimm_k_layernorm(ptr(x), H, rows, H, ptr(g), ptr(b), ptr(y32), ptr(y16), 0, stream()))
    yh = torch.empty(rows, H, device=DEV, dtype=torch.float16)
    check(lib.unimm_k_layernorm(ptr(x), H, rows, H, ptr(g), ptr(b), None, ptr(yh), 1, stream()))
    ref = torch.nn.functional.layer_norm(x.double(), (H,), g.double(), b.double(), 1e-12)
    assert (y32.double() - ref).abs().max().item() < 2e-5
    assert (y16.double() - ref).abs().max().item() < 0.03
    assert (yh.double() - ref).abs().max().item() < 0.004


# ----------------------------------------------------------------------------------------------- attention
def ref_attention(q, k, v, heads, allow):
    """softmax(QK^T/sqrt(d) + (1-allow)*-10000) V in fp64, exactly the reference's additive form."""
    B, Sq, HD = q.shape
    d = HD // heads
    qh = q.double().view(B, Sq, heads, d).permute(0, 2, 1, 3)
    kh = k.double().view(B, -1, heads, d).permute(0, 2, 1, 3)
    vh = v.double().view(B, -1, heads, d).permute(0, 2, 1, 3)
    s = qh @ kh.transpose(-1, -2) / math.sqrt(d) + (1.0 - allow.double())[:, None] * -10000.0
    return (torch.softmax(s, -1) @ vh).permute(0, 2, 1, 3).reshape(B, Sq, HD)


def make_desc():
    # gen: (ctx, L, last) incl. T == 256 edge; dis rows
    rows = [(0, 239, 247, 8), (0, 239, 241, 2), (0, 30, 35, 5), (0, 240, 248, 8), (1, 0, 200, 0), (1, 0, 256, 0), (0, 2, 4, 2)]
    return torch.tensor(rows, dtype=torch.int32, device=DEV)


ELEM = {0: torch.float32, 1: torch.bfloat16, 2: torch.float16}
ATOL = {0: 2e-5, 1: 3e-2, 2: 4e-3}


@pytest.mark.parametrize("kind,impl", [(0, 0), (1, 0), (1, 1), (2, 0), (2, 1), (1, 2), (2, 2)])     # impl 2 = tcgen05 / TMEM
def test_text_self_attention(kind, impl):
    desc = make_desc()
    B, S, heads, d = desc.shape[0], 256, 12, 64
    H = heads * d
    qkv = rnd(B * S, 3 * H, seed=11)
    allow = dense_text_mask(desc, S)
    valid = allow.any(-1)                                  # padding rows are garbage-by-design in the reference
    qkv = qkv.to(ELEM[kind])
    out = torch.zeros(B * S, H, device=DEV, dtype=qkv.dtype)
    e = qkv.element_size()
    base = qkv.data_ptr()
    check(lib.unimm_k_attention(C.c_void_p(base), 3 * H, C.c_void_p(base + e * H), 3 * H, C.c_void_p(base + 2 * e * H), 3 * H,
                                ptr(out), H, B, heads, d, S, S, _lib.MASK_TEXT_SELF, ptr(desc), None, kind, impl, stream()))
    q3 = qkv.view(B, S, 3 * H)
    ref = ref_attention(q3[..., :H], q3[..., H:2 * H], q3[..., 2 * H:], heads, allow)
    err = (out.view(B, S, H).double() - ref)[valid].abs().max().item()
    print(f"text self-attention kind={kind} impl={impl}: max err on valid rows {err:.3e}")
    assert err < ATOL[kind]
    assert torch.isfinite(out.float()).all()


@pytest.mark.parametrize("kind,impl", [(0, 0), (1, 0), (1, 1), (2, 0), (2, 1)])
def test_cross_and_image_attention(kind, impl):
    desc = make_desc()
    B, S, R, heads, d = desc.shape[0], 256, 37, 8, 128
    H = heads * d
    dt, tol = ELEM[kind], ATOL[kind]
    qkv_t, qkv_v = rnd(B * S, 3 * H, seed=21).to(dt), rnd(B * R, 3 * H, seed=22).to(dt)
    e = qkv_t.element_size()
    img_mask = torch.ones(B, R, device=DEV)
    img_mask[1, 30:] = 0
    img_mask[2, 1:] = 0
    # text -> image (key vector mask)
    o1 = torch.zeros(B * S, H, device=DEV, dtype=dt)
    check(lib.unimm_k_attention(ptr(qkv_t), 3 * H, C.c_void_p(qkv_v.data_ptr() + e * H), 3 * H,
                                C.c_void_p(qkv_v.data_ptr() + 2 * e * H), 3 * H, ptr(o1), H, B, heads, d, S, R,
                                _lib.MASK_KEY_VECTOR, None, ptr(img_mask), kind, impl, stream()))
    t3, v3 = qkv_t.view(B, S, 3 * H), qkv_v.view(B, R, 3 * H)
    ref1 = ref_attention(t3[..., :H], v3[..., H:2 * H], v3[..., 2 * H:], heads, img_mask[:, None, :].expand(B, S, R))
    err1 = (o1.view(B, S, H).double() - ref1).abs().max().item()
    # image -> text (co interval)
    o2 = torch.zeros(B * R, H, device=DEV, dtype=dt)
    check(lib.unimm_k_attention(ptr(qkv_v), 3 * H, C.c_void_p(qkv_t.data_ptr() + e * H), 3 * H,
                                C.c_void_p(qkv_t.data_ptr() + 2 * e * H), 3 * H, ptr(o2), H, B, heads, d, R, S,
                                _lib.MASK_CO_INTERVAL, ptr(desc), None, kind, impl, stream()))
    co = dense_co_mask(desc, S).bool()
    ref2 = ref_attention(v3[..., :H], t3[..., H:2 * H], t3[..., 2 * H:], heads, co[:, None, :].expand(B, R, S))
    err2 = (o2.view(B, R, H).double() - ref2).abs().max().item()
    # image self-attention
    o3 = torch.zeros(B * R, H, device=DEV, dtype=dt)
    check(lib.unimm_k_attention(ptr(qkv_v), 3 * H, C.c_void_p(qkv_v.data_ptr() + e * H), 3 * H,
                                C.c_void_p(qkv_v.data_ptr() + 2 * e * H), 3 * H, ptr(o3), H, B, heads, d, R, R,
                                _lib.MASK_KEY_VECTOR, None, ptr(img_mask), kind, impl, stream()))
    ref3 = ref_attention(v3[..., :H], v3[..., H:2 * H], v3[..., 2 * H:], heads, img_mask[:, None, :].expand(B, R, R))
    err3 = (o3.view(B, R, H).double() - ref3).abs().max().item()
    print(f"cross attention kind={kind} impl={impl}: t->i {err1:.3e}  i->t {err2:.3e}  i self {err3:.3e}")
    assert max(err1, err2, err3) < tol


def packed_candidate_layout(units, seed=0):
    """Synthetic prefix-shared layout: ``units`` = [(ctx_len, [last_len per candidate])].  Returns row count, candidate jobs,
    row_iv and, per candidate row, the list of allowed key rows (context + own interval + self) for the reference."""
    n_shared = sum(c for c, _ in units)
    jobs, iv, allowed = [], {}, {}
    s0, row = 0, n_shared
    for ctx, lasts in units:
        q_start = row
        for last in lasts:
            s_abs = row
            for idx in range(1 + 2 * last):
                r = s_abs + idx
                if idx == 0:
                    lo, hi, self_ = s_abs, s_abs + 1 + 2 * last, -1
                elif idx <= last:
                    lo, hi, self_ = s_abs + 1, s_abs + idx + 1, -1
                else:
                    lo, hi, self_ = s_abs + 1, s_abs + idx - last, r
                iv[r] = (lo, hi, self_, 0)
                allowed[r] = list(range(s0, s0 + ctx)) + list(range(lo, hi)) + ([self_] if self_ >= 0 else [])
            row += 1 + 2 * last
        jobs.append((q_start, row - q_start, s0, ctx, 1, -1, 0, 0))
        s0 += ctx
    M = row
    row_iv = torch.zeros(M, 4, dtype=torch.int32)
    row_iv[:, 2] = -1
    for r, v in iv.items():
        row_iv[r] = torch.tensor(v, dtype=torch.int32)
    return M, torch.tensor(jobs, dtype=torch.int32), row_iv, allowed


@pytest.mark.parametrize("lp", ["fp16", "bf16"])
@pytest.mark.parametrize("impl", [1, 2])
def test_candidate_attention_kernels(impl, lp):
    """Candidate rows over context U own rows: persistent mma.sync kernel (1) and tcgen05/TMEM kernel (2) against fp64."""
    g = np.random.RandomState(5)
    units = [(239, list(g.randint(1, 9, size=100))), (30, list(g.randint(1, 9, size=40))), (64, [8, 8, 1]), (129, list(g.randint(1, 9, size=57))),
             (256, list(g.randint(1, 9, size=9))), (1, [3])]
    M, jobs, row_iv, allowed = packed_candidate_layout(units)
    heads, d = 12, 64
    H = heads * d
    dt = torch.float16 if lp == "fp16" else torch.bfloat16
    qkv = rnd(M + 7, 3 * H, seed=31).to(dt)[:M]            # a few allocated rows behind the tensor extent
    out = torch.full((M, H), float("nan"), device=DEV, dtype=dt)
    e, base = qkv.element_size(), qkv.data_ptr()
    max_q = int(jobs[:, 1].max())
    dj, di = jobs.to(DEV), row_iv.to(DEV)
    check(lib.unimm_k_attention_jobs(C.c_void_p(base), 3 * H, C.c_void_p(base + e * H), 3 * H, C.c_void_p(base + 2 * e * H), 3 * H,
                                     ptr(out), H, M, heads, d, ptr(dj), jobs.shape[0], max_q, 256, 192, ptr(di), 16,
                                     1 if lp == "fp16" else 0, impl, stream()))
    torch.cuda.synchronize()
    q = qkv[:, :H].double().view(M, heads, d)
    k = qkv[:, H:2 * H].double().view(M, heads, d)
    v = qkv[:, 2 * H:].double().view(M, heads, d)
    rows = sorted(allowed)
    err = 0.0
    for r in rows[::7] + rows[-40:]:
        keys = torch.tensor(allowed[r], device=DEV)
        s = torch.einsum("hd,khd->hk", q[r], k[keys]) / math.sqrt(d)
        ref = torch.einsum("hk,khd->hd", torch.softmax(s, -1), v[keys]).reshape(H)
        err = max(err, (out[r].double() - ref).abs().max().item())
    n_shared = sum(c for c, _ in units)
    assert torch.isfinite(out[n_shared:].float()).all()
    assert torch.isnan(out[:n_shared].float()).all()       # context rows are not this kernel's to write
    print(f"candidate attention impl={impl} {lp}: max err {err:.3e}")
    assert err < (4e-3 if lp == "fp16" else 3e-2)


@pytest.mark.parametrize("lp", ["fp16", "bf16"])
@pytest.mark.parametrize("impl", [0, 2])
def test_text_to_image_jobs(impl, lp):
    """Packed text rows over their unit's 37 image regions (D = 128, image padding mask): generic job kernel and tcgen05 kernel."""
    R, heads, d = 37, 8, 128
    H = heads * d
    q_lens = [239, 1100, 30, 700, 1, 129, 256, 64]            # two jobs per unit (context rows, candidate rows)
    U = len(q_lens) // 2
    jobs, row = [], 0
    for i, n in enumerate(q_lens):
        jobs.append((row, n, (i // 2) * R, R, 0, i // 2, 0, 0))
        row += n
    Mt, Mv = row, U * R
    dt = torch.float16 if lp == "fp16" else torch.bfloat16
    qkv_t, qkv_v = rnd(Mt, 3 * H, seed=41).to(dt), rnd(Mv, 3 * H, seed=42).to(dt)
    mask = torch.ones(U, R, device=DEV)
    mask[1, 30:] = 0
    mask[2, 1:] = 0
    mask[3, :] = 0                                            # nothing valid: every key allowed (the reference's additive mask)
    out = torch.full((Mt, H), float("nan"), device=DEV, dtype=dt)
    dj = torch.tensor(jobs, dtype=torch.int32, device=DEV)
    e = qkv_v.element_size()
    check(lib.unimm_k_attention_cross_jobs(ptr(qkv_t), 3 * H, C.c_void_p(qkv_v.data_ptr() + e * H), 3 * H,
                                           C.c_void_p(qkv_v.data_ptr() + 2 * e * H), 3 * H, ptr(out), H, Mt, Mv, heads, d, ptr(dj),
                                           len(jobs), max(q_lens), ptr(mask), R, 1 if lp == "fp16" else 0, impl, stream()))
    torch.cuda.synchronize()
    err = 0.0
    for (q0, n, k0, _, _, u, _, _) in jobs:
        q = qkv_t[q0:q0 + n, :H].double().view(n, heads, d)
        k = qkv_v[k0:k0 + R, H:2 * H].double().view(R, heads, d)
        v = qkv_v[k0:k0 + R, 2 * H:].double().view(R, heads, d)
        sc = torch.einsum("qhd,khd->hqk", q, k) / math.sqrt(d)
        m = mask[u].bool() if mask[u].any() else torch.ones(R, dtype=torch.bool, device=DEV)
        sc = sc.masked_fill(~m[None, None, :], float("-inf"))
        ref = torch.einsum("hqk,khd->qhd", torch.softmax(sc, -1), v).reshape(n, H)
        err = max(err, (out[q0:q0 + n].double() - ref).abs().max().item())
    assert torch.isfinite(out.float()).all()
    print(f"text->image jobs impl={impl} {lp}: max err {err:.3e}")
    assert err < (4e-3 if lp == "fp16" else 3e-2)


def test_verify_masks_kernel():
    desc = make_desc()
    B, S, R = desc.shape[0], 256, 37
    txt = dense_text_mask(desc, S).contiguous()
    co = dense_co_mask(desc, S).unsqueeze(1).repeat(1, R, 1).contiguous()
    flag = torch.zeros(1, dtype=torch.int32, device=DEV)
    check(lib.unimm_verify_masks(ptr(desc), B, S, R, ptr(txt.to(torch.uint8)), 1, ptr(co), ptr(flag), stream()))
    assert flag.item() == 0
    txt2 = txt.clone()
    txt2[3, 100, 5] = ~txt2[3, 100, 5]
    check(lib.unimm_verify_masks(ptr(desc), B, S, R, ptr(txt2.long().contiguous()), 8, ptr(co), ptr(flag), stream()))
    assert flag.item() == 1
