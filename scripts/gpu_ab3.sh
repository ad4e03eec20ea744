# A/B/C of UNIMM_GEMM_MULTICAST modes on the packed bench
mkdir -p gpurun_out
for rep in 1 2; do for v in 2 1; do
  UNIMM_GEMM_MULTICAST=$v timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ab.log 2>&1; tail -1 gpurun_out/bench_ab.log | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('pair mode $v cand/s', round(d['value']), 'ms', round(d['ms_per_step'],2), 'gemm TF', round(d['roofline']['achieved']), d['roofline']['share_of_step'], d['clocks']['sm_mhz'])"
done; done
