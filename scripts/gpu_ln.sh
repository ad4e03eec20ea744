mkdir -p gpurun_out
P="python -m pytest -q -s -p no:cacheprovider"
UNIMM_DEBUG=1 timeout 300 $P tests/test_kernels_gpu.py -x -k "gemm_ln" > gpurun_out/k_ln.log 2>&1; tail -2 gpurun_out/k_ln.log; grep -E "FAIL|rror|co-resident" gpurun_out/k_ln.log | sort | uniq -c | head -8
for mc in 1 0; do echo "== LN pair $mc"; UNIMM_LN_PAIR=$mc timeout 300 python scripts/gemm_bench.py 98176 2>&1 | tail -8 | grep 98176 | grep -v img_out | cut -c100-250; done
for rep in 1 2; do for v in 1 0; do
  UNIMM_LN_PAIR=$v timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ab.log 2>&1; tail -1 gpurun_out/bench_ab.log | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('LN pair $v cand/s', round(d['value']), 'ms', round(d['ms_per_step'],2), 'gemm TF', round(d['roofline']['achieved']), round(d['roofline']['other_tensor_kernels_tflops']['umma_gemm_ln_kernel (LayerNorm-fused cluster GEMM)']), d['roofline']['share_of_step'], d['clocks']['sm_mhz'])"
done; done
