mkdir -p gpurun_out
timeout 600 python bench.py --steps 20 --warmup 3 --cpu-sample 25 > gpurun_out/bench_ab.log 2>&1; tail -1 gpurun_out/bench_ab.log | python -c "
import sys,json
d=json.loads(sys.stdin.read()); r=d['roofline']; print(round(d['value']), round(d['e2e']['value']), r['achieved'], r['traffic'], r['algorithmic_bytes'], r['traffic_source'], d['clocks'])"
for m in tf32 bf16; do
  timeout 300 python bench.py --impl reference --ref-device cuda --ref-mode $m --steps 3 --warmup 1 > gpurun_out/bench_eager_$m.log 2>&1; tail -1 gpurun_out/bench_eager_$m.log | cut -c1-330
done
