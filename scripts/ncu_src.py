#!/usr/bin/env python
"""Per-opcode / per-source-line stall summary of one kernel from an ncu report (`--page source --csv`).
usage: python scripts/ncu_src.py <report.ncu-rep> <kernel regex> [top-N lines]"""
import collections, csv, subprocess, sys
rep, pat = sys.argv[1], sys.argv[2]
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{pat}", "--launch-count", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = next(r for r in rows if "# Samples" in r)
idx = {h: i for i, h in enumerate(hdr)}
def num(r, h):
    try: return int(r[idx[h]])
    except Exception: return 0
data = [r for r in rows if len(r) == len(hdr) and r[idx["# Samples"]].isdigit()]
tot = sum(num(r, "# Samples") for r in data) or 1
agg, inst = collections.Counter(), collections.Counter()
for r in data:
    p = r[idx["Source"]].strip().split()
    op = p[0] if p else ""
    if op.startswith("@") and len(p) > 1: op = p[1]
    op = ".".join(op.split(".")[:2])
    agg[op] += num(r, "# Samples"); inst[op] += num(r, "Instructions Executed")
ti = sum(inst.values()) or 1
print(f"# {pat}: {tot} samples, {ti} warp instructions")
for k, v in agg.most_common(topn): print(f"{k:22s} samples {100*v/tot:5.1f}%  inst {100*inst[k]/ti:5.1f}%")
st = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
s = collections.Counter({h: sum(num(r, h) for r in data) for h in st})
print("stalls:", {k: round(100 * v / tot, 1) for k, v in s.most_common(8)})
print("# hottest instructions")
for r in sorted(data, key=lambda r: -num(r, "# Samples"))[:topn]:
    tops = sorted(((num(r, h), h) for h in st), reverse=True)[:2]
    print(f"{100*num(r,'# Samples')/tot:5.1f}%  {r[idx['Source']].strip()[:90]:90s} {tops[0][1]}={tops[0][0]} {tops[1][1]}={tops[1][0]}")
