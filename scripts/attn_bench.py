#!/usr/bin/env python
"""Microbenchmark of the candidate-row attention kernels on a bench-sized packed layout (80 units x 100 candidates).
usage: python scripts/attn_bench.py [impl ...]    (1 = persistent mma.sync kernel, 2 = tcgen05 kernel)"""
import ctypes as C, math, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from unimm_b200._lib import check, lib, ptr
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))

def layout(units):
    n_shared = sum(c for c, _ in units)
    jobs, rows_iv = [], []
    s0, row = 0, n_shared
    for ctx, lasts in units:
        q_start = row
        for last in lasts:
            s_abs = row
            for idx in range(1 + 2 * last):
                r = s_abs + idx
                if idx == 0: rows_iv.append((s_abs, s_abs + 1 + 2 * last, -1, 0))
                elif idx <= last: rows_iv.append((s_abs + 1, s_abs + idx + 1, -1, 0))
                else: rows_iv.append((s_abs + 1, s_abs + idx - last, r, 0))
            row += 1 + 2 * last
        jobs.append((q_start, row - q_start, s0, ctx, 1, -1, 0, 0))
        s0 += ctx
    iv = np.zeros((row, 4), np.int32); iv[:, 2] = -1
    iv[n_shared:] = np.asarray(rows_iv, np.int32)
    return row, np.asarray(jobs, np.int32), iv

dev = torch.device("cuda", 0)
g = np.random.RandomState(0)
units = []
for img in range(8):
    for rnd in range(10):
        ctx = min(255, 22 + rnd * 22 + int(g.randint(0, 8)))
        units.append((ctx, list(g.randint(2, 9, size=100))))
M, jobs, iv = layout(units)
heads, d = 12, 64
H = heads * d
qkv = (torch.randn(M, 3 * H, device=dev)).half()
out = torch.zeros(M, H, device=dev, dtype=torch.float16)
dj, di = torch.from_numpy(jobs).to(dev), torch.from_numpy(iv).to(dev)
st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
e, base = 2, qkv.data_ptr()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
pairs = sum(sum((1 + 2 * l) * c for l in ls) for c, ls in units)
print(f"rows {M}, units {len(units)}, ctx pairs {pairs/1e6:.1f} M, ctx FLOPs {4*heads*d*pairs/1e9:.1f} G, Q+O+K+V own bytes {4*(M-sum(c for c,_ in units))*H*2/1e6:.0f} MB")
ref = None
FLUSH = os.environ.get("ATTN_BENCH_FLUSH", "1") != "0"
for impl in [int(x) for x in (sys.argv[1:] or ["1", "2"])]:
    def run():
        check(lib.unimm_k_attention_jobs(C.c_void_p(base), 3 * H, C.c_void_p(base + e * H), 3 * H, C.c_void_p(base + 2 * e * H), 3 * H,
                                         ptr(out), H, M, heads, d, ptr(dj), jobs.shape[0], int(jobs[:, 1].max()), 256, 192, ptr(di), 16, 1, impl, st))
    for _ in range(3): run()
    ts = []
    for _ in range(10):
        if FLUSH: flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); run(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    o = out.float().clone()
    if ref is None: ref = o
    print(f"impl {impl}: median {np.median(ts):.1f} us  min {min(ts):.1f} us   max diff vs first impl {(o - ref).abs().max().item():.2e}")
