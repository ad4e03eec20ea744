mkdir -p gpurun_out
for m in "packed 8" "packed 16" "packed 2"; do
  set -- $m
  timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --mode $1 --images-per-step $2 > gpurun_out/bench_$1_$2.log 2>&1
  python - <<PY
import json
l=[x for x in open("gpurun_out/bench_$1_$2.log") if x.startswith("{")]
if not l: print(open("gpurun_out/bench_$1_$2.log").read()[-3000:])
else:
    d=json.loads(l[-1]); print("$1 $2:", round(d["value"]), "e2e", round(d["e2e"]["value"]), "gemm TF", round(d["roofline"]["achieved"]), d["roofline"]["share_of_step"], {k:(round(v,3) if isinstance(v,float) else v) for k,v in d["pct_of_bf16_peak"].items()}, "ms/step", round(d["ms_per_step"],1), d["config"].get("packed_text_rows_per_step"), d["clocks"])
PY
done
