mkdir -p gpurun_out
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ab.log 2>&1; tail -1 gpurun_out/bench_ab.log | python -c "
import sys,json
d=json.loads(sys.stdin.read()); r=d['roofline']; print(round(d['value']), round(d['e2e']['value']), round(r['achieved']), r['traffic_source'], d['clocks'])" || tail -5 gpurun_out/bench_ab.log
