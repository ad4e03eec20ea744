#!/usr/bin/env python
"""Turn the ncu artefacts gpurun brings back (gpurun_out/) into the small text summaries kept under profiles/.

usage: python scripts/summarize_ncu.py <tag>        e.g. r01_v1
reads  gpurun_out/launches.csv          (ncu --metrics gpu__time_duration.sum ... --csv)
       gpurun_out/prof_*.ncu-rep        (ncu --set full), via `ncu -i ... --page raw --csv`
"""
import collections
import csv
import glob
import os
import re
import subprocess
import sys

tag = sys.argv[1]
os.makedirs("profiles", exist_ok=True)


def short(name):
    name = re.sub(r"^void ", "", name)
    name = name.replace("unimm::<unnamed>::", "").replace("unnamed>::", "")
    return re.sub(r"\(.*", "", name)


if os.path.exists("gpurun_out/launches.csv"):
    lines = [l for l in open("gpurun_out/launches.csv") if not l.startswith("==")]
    agg, tot, seq = collections.defaultdict(lambda: [0, 0.0]), 0.0, []
    for row in csv.DictReader(lines):
        try:
            v = float(row["Metric Value"].replace(",", ""))
        except ValueError:
            continue
        v = v / 1e3 if row["Metric Unit"] == "ns" else (v * 1e3 if row["Metric Unit"] == "ms" else v)
        k = short(row["Kernel Name"])
        agg[k][0] += 1
        agg[k][1] += v
        tot += v
        seq.append((k, v))
    with open(f"profiles/{tag}_launches.txt", "w") as f:
        f.write("# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare SHARES)\n")
        f.write(f"# command: see scripts/gpu_profile.sh; total {tot:.0f} us over {len(seq)} launches\n")
        f.write(f"{'us':>12} {'share':>7} {'n':>6} {'avg us':>9}  kernel\n")
        for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:30]:
            f.write(f"{t:12.1f} {100 * t / tot:6.1f}% {n:6d} {t / n:9.2f}  {k[:110]}\n")
        f.write("\n# launch order (our kernels only), one text-layer period and one connection-layer period:\n")
        ours = [(k, v) for k, v in seq if not k.startswith("at")]
        f.write(" ".join(f"{k.split('<')[0].replace('_kernel', '')}:{v:.0f}" for k, v in ours[:140]) + "\n")
    print("wrote", f"profiles/{tag}_launches.txt")

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__cycles_elapsed.max",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
for rep in sorted(glob.glob("gpurun_out/prof_*.ncu-rep")):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    if len(rows) < 3:
        continue
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    stall = [h for h in hdr if "warp_issue_stalled" in h and h.endswith("per_warp_active.pct")]
    name = os.path.basename(rep).replace(".ncu-rep", "")
    with open(f"profiles/{tag}_{name}.txt", "w") as f:
        f.write(f"# ncu --set full --clock-control none --import-source on ({os.path.basename(rep)}); one block per captured launch\n")
        for r in rows[2:]:
            f.write(f"\n== {short(r[idx['Kernel Name']])}  grid {r[idx['Grid Size']]} block {r[idx['Block Size']]}\n")
            for k in KEYS:
                if k in idx:
                    f.write(f"  {k} [{units[idx[k]]}] = {r[idx[k]]}\n")
            top = sorted(((float(r[idx[h]]), h) for h in stall if r[idx[h]] not in ("", "n/a")), reverse=True)[:6]
            for v, h in top:
                f.write(f"  stall {h.split('stalled_')[1].split('_per_warp')[0]} = {v:.1f} % of warp-active cycles\n")
    print("wrote", f"profiles/{tag}_{name}.txt")
