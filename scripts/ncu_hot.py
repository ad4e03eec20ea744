#!/usr/bin/env python
"""Top stall lines of an `ncu --page source --csv` dump (SASS view): python scripts/ncu_hot.py src.csv [top]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = rows[1]
col = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
body = rows[2:]
tot = sum(int(r[col["# Samples"]] or 0) for r in body)
print("total samples", tot)
agg = {}
for h in stall_cols:
    agg[h] = sum(int(r[col[h]] or 0) for r in body)
print("by reason:", {k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v > tot * 0.01})
idx = sorted(range(len(body)), key=lambda i: -int(body[i][col["# Samples"]] or 0))[:top]
for i in sorted(idx):
    r = body[i]
    reasons = {h[6:]: int(r[col[h]] or 0) for h in stall_cols if int(r[col[h]] or 0) > 0}
    main = sorted(reasons.items(), key=lambda kv: -kv[1])[:3]
    print(f"{i:5d} {int(r[col['# Samples']]):7d} {100*int(r[col['# Samples']])/tot:5.1f}%  exec {r[col['Instructions Executed']]:>8s}  {r[col['Source']].strip()[:90]:90s} {main}")
