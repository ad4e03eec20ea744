mkdir -p gpurun_out
P="python -m pytest -q -s -p no:cacheprovider"
UNIMM_LN_RES=precharge timeout 900 $P tests/test_kernels_gpu.py -k "gemm_ln" > gpurun_out/k_ln.log 2>&1; tail -2 gpurun_out/k_ln.log; grep -E "FAIL|rror" gpurun_out/k_ln.log | head -5
UNIMM_LN_RES=precharge timeout 900 $P tests/test_parity_gpu.py -k "fp16" > gpurun_out/p_ln.log 2>&1; tail -2 gpurun_out/p_ln.log; grep -E "FAIL|rror" gpurun_out/p_ln.log | head -5
for mc in precharge mma; do echo "== LN residual $mc"; UNIMM_LN_RES=$mc timeout 600 python scripts/gemm_bench.py 98176 2>&1 | tail -8 | grep 98176 | grep -v img_out | cut -c100-250; done
for rep in 1 2; do for v in precharge mma; do
  UNIMM_LN_RES=$v timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ab.log 2>&1; tail -1 gpurun_out/bench_ab.log | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('LN residual $v cand/s', round(d['value']), 'ms', round(d['ms_per_step'],2), 'gemm TF', round(d['roofline']['achieved']), d['roofline']['share_of_step'], d['clocks']['sm_mhz'])"
done; done
