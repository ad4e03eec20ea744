mkdir -p gpurun_out
P="python -m pytest -q -s -p no:cacheprovider"
timeout 900 $P tests/test_kernels_gpu.py -k "gemm_ln" > gpurun_out/k_ln.log 2>&1; tail -2 gpurun_out/k_ln.log; grep -E "FAIL|rror" gpurun_out/k_ln.log | head -5
for mc in 4 3; do echo "== LN split $mc"; UNIMM_LN_SPLIT=$mc timeout 600 python scripts/gemm_bench.py 98176 2>&1 | tail -8 | grep 98176 | grep -v img_out | cut -c1-250; done
for rep in 1 2; do for v in 4 3; do
  UNIMM_LN_SPLIT=$v timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ab.log 2>&1; tail -1 gpurun_out/bench_ab.log | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('LN split $v cand/s', round(d['value']), 'ms', round(d['ms_per_step'],2), 'gemm TF', round(d['roofline']['achieved']), d['roofline']['share_of_step'], d['clocks']['sm_mhz'])"
done; done
