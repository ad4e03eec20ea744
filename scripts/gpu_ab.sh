# A/B of an environment switch on the packed bench: usage VAR=UNIMM_GEMM_MULTICAST bash scripts/gpu_ab.sh
mkdir -p gpurun_out
for rep in ${REPS:-1 2}; do for v in 1 0; do
  env ${VAR}=$v timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ab.log 2>&1; tail -1 gpurun_out/bench_ab.log | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('${VAR}=$v cand/s', round(d['value']), 'ms', round(d['ms_per_step'],2), 'gemm TF', round(d['roofline']['achieved']), d['roofline']['share_of_step'], d['clocks']['sm_mhz'])"
done; done
