#!/usr/bin/env python
"""Per-kernel counts of the SASS mnemonics that prove tcgen05 / TMEM / TMA use (B200_PROFILING.md): UTCHMMA (tcgen05.mma),
LDTM / STTM (tcgen05.ld / st), UTMALDG (TMA tensor load), UTCBAR (tcgen05.commit), HMMA (mma.sync), from
`cuobjdump -sass unimm_b200/lib/libunimm_b200.so`.   python scripts/sass_summary.py > profiles/r02_sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "unimm_b200", "lib", "libunimm_b200.so")
MNEMONICS = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTCBAR", "HMMA", "SYNCS", "MUFU.EX2", "FFMA2"]

out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
counts, order, cur = {}, [], None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = cur.replace("unimm::(anonymous namespace)::", "").replace("(anonymous namespace)::", "").replace("void ", "")
        cur = re.sub(r"\(.*", "", cur)
        counts[cur] = collections.Counter()
        order.append(cur)
        continue
    if cur is None:
        continue
    for mn in MNEMONICS:
        if re.search(r"\b" + re.escape(mn) + r"\b", line) or (mn + ".") in line:
            counts[cur][mn] += 1
print("kernel".ljust(72), " ".join(m.rjust(8) for m in MNEMONICS))
tot = collections.Counter()
for k in order:
    c = counts[k]
    if not any(c[m] for m in MNEMONICS[:8]):
        continue
    print(k[:72].ljust(72), " ".join(str(c[m]).rjust(8) for m in MNEMONICS))
    tot.update(c)
print("TOTAL".ljust(72), " ".join(str(tot[m]).rjust(8) for m in MNEMONICS))
