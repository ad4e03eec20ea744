#!/usr/bin/env python
"""Runs the LayerNorm-fused cluster GEMM a few times on one shape (for ncu).  Usage: python scripts/ln_probe.py [N K M]"""
import ctypes as C
import sys

import torch

sys.path.insert(0, ".")
from unimm_b200._lib import check, lib, ptr  # noqa: E402

dev = torch.device("cuda", 0)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 768
K = int(sys.argv[2]) if len(sys.argv) > 2 else 768
m = int(sys.argv[3]) if len(sys.argv) > 3 else 64000
st = lambda: C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
A = (torch.randn(m, K, device=dev) * 0.5).half()
W = (torch.randn(N, K, device=dev) * 0.05).half()
Wp = torch.empty_like(W)
check(lib.unimm_k_permute_w_ln(ptr(W), ptr(Wp), N, K, st()))
bias, gamma, beta = torch.randn(N, device=dev), torch.rand(N, device=dev) + 0.5, torch.randn(N, device=dev)
X = torch.randn(m, N, device=dev)
o16 = torch.empty(m, N, device=dev, dtype=torch.float16)
X16 = X.half()
for _ in range(4):
    if len(sys.argv) > 4 and sys.argv[4] == "fp32res":
        check(lib.unimm_k_gemm_ln_lp(ptr(A), K, ptr(Wp), K, m, N, K, ptr(bias), ptr(X), N, None, 0, ptr(gamma), ptr(beta), ptr(X), N, ptr(o16), N, 1, st()))
    else:
        check(lib.unimm_k_gemm_ln_lp(ptr(A), K, ptr(Wp), K, m, N, K, ptr(bias), None, 0, ptr(X16), N, ptr(gamma), ptr(beta), None, 0, ptr(X16), N, 1, st()))
torch.cuda.synchronize()
print("ok")
