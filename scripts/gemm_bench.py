#!/usr/bin/env python
"""Per-shape timing of the tcgen05 GEMM through the C ABI (CUDA events, L2 flushed between launches),
next to torch.matmul (cuBLAS) on the same shapes.  Usage: python scripts/gemm_bench.py [M]"""
import ctypes as C
import sys

import torch

sys.path.insert(0, ".")
from unimm_b200._lib import check, lib, ptr  # noqa: E402

dev = torch.device("cuda", 0)
M = int(sys.argv[1]) if len(sys.argv) > 1 else 64000
MODES = len(sys.argv) > 2 and sys.argv[2] == "modes"     # also time the epilogue ablations: 1 = row-per-thread stores, 2 = no stores, 3 = no epilogue
st = lambda: C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

# (name, N, K, act, residual+fp32 out, 16-bit out)
SHAPES = [("qkv", 2304, 768, 0, False, True), ("out_proj", 768, 768, 0, True, False), ("ffn1_gelu", 3072, 768, 1, False, True),
          ("ffn2", 768, 3072, 0, True, False), ("bi_qkv_t", 3072, 768, 0, False, True), ("bi_dense2", 768, 1024, 0, True, False),
          ("plain_4096", 4096, 4096, 0, False, True)]


def timeit(fn, iters=8):
    ts = []
    for _ in range(iters + 2):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts = sorted(ts[2:])
    return ts[len(ts) // 2]


for name, N, K, act, res, lp_out in SHAPES:
    m = M if name != "plain_4096" else 16384
    A = (torch.randn(m, K, device=dev) * 0.5).half()
    W = (torch.randn(N, K, device=dev) * 0.05).half()
    bias = torch.randn(N, device=dev)
    R = torch.randn(m, N, device=dev) if res else None
    o32 = torch.empty(m, N, device=dev) if res else None
    o16 = torch.empty(m, N, device=dev, dtype=torch.float16) if lp_out else None
    fn = lambda: check(lib.unimm_k_gemm_lp(ptr(A), K, ptr(W), K, m, N, K, ptr(bias), ptr(R), N, act, ptr(o32), N, ptr(o16), N, 0, 0, 1, st()))
    t = timeit(fn)
    t_ref = timeit(lambda: torch.matmul(A, W.t()))
    fl = 2.0 * m * N * K
    extra = ""
    if lp_out and not res:    # fragment-ordered weights: smem-free epilogue (+ the tanh-form GELU)
        Wp = torch.empty_like(W)
        check(lib.unimm_k_permute_w(ptr(W), ptr(Wp), N, K, 1, st()))
        ff = lambda a=act: check(lib.unimm_k_gemm_lp(ptr(A), K, ptr(Wp), K, m, N, K, ptr(bias), None, 0, a, None, 0, ptr(o16), N, 0, 0, 1 | 0x100, st()))
        extra += f" | frag {fl/timeit(ff)/1e9:6.0f}"
        if act == 1:
            extra += f" | frag+tanh {fl/timeit(lambda: ff(3))/1e9:6.0f}"
    if MODES:
        for mode in (1, 2, 3, 4, 7):
            fm = lambda: check(lib.unimm_k_gemm_lp(ptr(A), K, ptr(W), K, m, N, K, ptr(bias), ptr(R), N, act, ptr(o32), N, ptr(o16), N,
                                                    1000 * mode, 0, 1, st()))
            extra += f" | mode{mode} {fl/timeit(fm)/1e9:6.0f}"
    print(f"{name:12s} M={m:6d} N={N:5d} K={K:5d} act={act} res={int(res)}: ours {t*1e3:8.1f} us {fl/t/1e9:7.1f} TFLOP/s | "
          f"cuBLAS (no epilogue) {t_ref*1e3:8.1f} us {fl/t_ref/1e9:7.1f} TFLOP/s{extra}", flush=True)


# LayerNorm-fused cluster GEMM against the unfused pair (GEMM with fp32 residual epilogue + LayerNorm kernel)
print("# fused = one cluster kernel; unfused = umma GEMM (+residual, fp32 out) followed by layernorm_rows", flush=True)
for name, N, K in (("out_proj+ln", 768, 768), ("ffn2+ln", 768, 3072), ("bi_dense2+ln", 768, 1024), ("img_out+ln", 1024, 1024)):
    for m in (M, 98176):
        A = (torch.randn(m, K, device=dev) * 0.5).half()
        W = (torch.randn(N, K, device=dev) * 0.05).half()
        bias, gamma, beta = torch.randn(N, device=dev), torch.rand(N, device=dev) + 0.5, torch.randn(N, device=dev)
        X = torch.randn(m, N, device=dev)
        pre = torch.empty(m, N, device=dev)
        o16 = torch.empty(m, N, device=dev, dtype=torch.float16)
        Wp = torch.empty_like(W)
        check(lib.unimm_k_permute_w_ln(ptr(W), ptr(Wp), N, K, st()))
        X16 = X.half()
        fused = lambda: check(lib.unimm_k_gemm_ln_lp(ptr(A), K, ptr(Wp), K, m, N, K, ptr(bias), ptr(X), N, None, 0, ptr(gamma), ptr(beta), ptr(X), N,
                                                     ptr(o16), N, 1, st()))
        fused16 = lambda: check(lib.unimm_k_gemm_ln_lp(ptr(A), K, ptr(Wp), K, m, N, K, ptr(bias), None, 0, ptr(X16), N, ptr(gamma), ptr(beta), None, 0,
                                                       ptr(X16), N, 1, st()))

        def unfused():
            check(lib.unimm_k_gemm_lp(ptr(A), K, ptr(W), K, m, N, K, ptr(bias), ptr(X), N, 0, ptr(pre), N, None, 0, 0, 0, 1, st()))
            check(lib.unimm_k_layernorm(ptr(pre), N, m, N, ptr(gamma), ptr(beta), ptr(X), ptr(o16), 1, st()))
        tf, t16, tu = timeit(fused), timeit(fused16), timeit(unfused)
        fl = 2.0 * m * N * K
        hbm = m * N * (4 + 4 + 2) + m * K * 2
        hbm16 = m * N * (2 + 2) + m * K * 2
        print(f"{name:13s} M={m:6d} N={N:5d} K={K:5d}: fused/fp32-residual {tf*1e3:7.1f} us {fl/tf/1e9:6.1f} TFLOP/s {hbm/tf/1e6:5.0f} GB/s | "
              f"fused/16-bit-residual {t16*1e3:7.1f} us {fl/t16/1e9:6.1f} TFLOP/s {hbm16/t16/1e6:5.0f} GB/s | unfused {tu*1e3:7.1f} us", flush=True)
