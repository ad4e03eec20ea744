#!/usr/bin/env python
"""Microbenchmark of the dense text self-attention kernels (B = 250 sequences of 256 rows): impl 1 = mma.sync, 2 = tcgen05."""
import ctypes as C, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from unimm_b200 import _lib
from unimm_b200._lib import check, lib, ptr
dev = torch.device("cuda", 0)
B, S, heads, d = 250, 256, 12, 64
H = heads * d
g = np.random.RandomState(0)
rows = []
for b in range(B):
    ctx = 30 + (b % 10) * 22
    last = int(g.randint(2, 9))
    rows.append((0, ctx, ctx + last, last))
desc = torch.tensor(rows, dtype=torch.int32, device=dev)
qkv = torch.randn(B * S, 3 * H, device=dev).half()
out = torch.zeros(B * S, H, device=dev, dtype=torch.float16)
st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
base, e = qkv.data_ptr(), 2
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
ref = None
for impl in [int(x) for x in (sys.argv[1:] or ["1", "2"])]:
    run = lambda: check(lib.unimm_k_attention(C.c_void_p(base), 3 * H, C.c_void_p(base + e * H), 3 * H, C.c_void_p(base + 2 * e * H), 3 * H,
                                              ptr(out), H, B, heads, d, S, S, _lib.MASK_TEXT_SELF, ptr(desc), None, 2, impl, st))
    for _ in range(3): run()
    ts = []
    for _ in range(10):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); run(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    o = out.float().clone()
    if ref is None: ref = o
    print(f"impl {impl}: median {np.median(ts):.1f} us  ({4*B*heads*S*S*d/np.median(ts)/1e6:.0f} TFLOP/s dense-equivalent)  max diff {(o-ref).abs().max().item():.2e}")
