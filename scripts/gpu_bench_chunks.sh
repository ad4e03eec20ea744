mkdir -p gpurun_out
for c in 500 1000; do
  timeout 900 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --chunk $c > gpurun_out/bench_c$c.log 2>&1
  python - <<PY
import json
l=[x for x in open("gpurun_out/bench_c$c.log") if x.startswith("{")]
if not l: print(open("gpurun_out/bench_c$c.log").read()[-2000:])
else:
    d=json.loads(l[-1]); print("chunk $c", round(d["value"]), "e2e", round(d["e2e"]["value"]), "gemm TF", round(d["roofline"]["achieved"]), d["roofline"]["share_of_step"], d["clocks"])
PY
done
