#!/usr/bin/env python
"""profiles/<tag>_gemm_traffic.json from an ncu pass over one bench step (see scripts/gpu_traffic.sh):
DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) per launch of umma_gemm_kernel (log-sum-exp instantiation excluded),
averaged over the launches of ONE step, next to the algorithmic bytes of the same launches.  usage: gemm_traffic.py <tag>"""
import collections, csv, json, sys
tag = sys.argv[1]
lines = [l for l in open("gpurun_out/traffic.csv") if not l.startswith("==")]
rows = collections.OrderedDict()
for r in csv.DictReader(lines):
    try: v = float(r["Metric Value"].replace(",", ""))
    except ValueError: continue
    u = r["Metric Unit"].lower()
    scale = {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "ns": 1e-3, "us": 1, "ms": 1e3}.get(u, 1)
    rows.setdefault(r["ID"], {"name": r["Kernel Name"]})[r["Metric Name"]] = v * scale
sel = [r for r in rows.values() if "umma_gemm_kernel" in r["name"] and "umma_gemm_ln" not in r["name"] and "umma_gemm_kernel<256, 1," not in r["name"]]
n = len(sel)
tot = sum(r["dram__bytes_read.sum"] + r["dram__bytes_write.sum"] for r in sel)
us = sum(r["gpu__time_duration.sum"] for r in sel)
out = {"kernel": "umma_gemm_kernel (all non-LSE instantiations of one bench step)", "launches_per_step": n, "bytes_per_launch": tot / n,
       "dram_read_bytes_per_step": sum(r["dram__bytes_read.sum"] for r in sel), "dram_write_bytes_per_step": sum(r["dram__bytes_write.sum"] for r in sel),
       "ncu_time_us_per_step": us,
       "how": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:umma_gemm_kernel "
              "over the second step of `bench.py --steps 1 --warmup 1` (packed, 8 images)"}
json.dump(out, open(f"profiles/{tag}_gemm_traffic.json", "w"), indent=1)
print(json.dumps(out, indent=1))
