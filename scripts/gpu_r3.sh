mkdir -p gpurun_out
P="python -m pytest -q -s -p no:cacheprovider"
timeout 600 $P tests/test_kernels_gpu.py -x -k "fragment" > gpurun_out/k_frag.log 2>&1; tail -2 gpurun_out/k_frag.log; grep -E "gemm_umma_frag\[|FAIL|rror" gpurun_out/k_frag.log | head -20
timeout 600 python scripts/gemm_bench.py 98176 nomodes > gpurun_out/gemm_bench3.log 2>&1; cat gpurun_out/gemm_bench3.log
timeout 1800 $P tests/test_parity_gpu.py > gpurun_out/p_all.log 2>&1; grep -E "passed|failed|FAIL" gpurun_out/p_all.log | cut -c1-200 | head; grep -E "^\.?\[fp16\]" gpurun_out/p_all.log | cut -c1-160
for v in "" "UNIMM_GELU_TANH=1" "UNIMM_FRAG_EPILOGUE=0"; do
  env $v timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_v.log 2>&1; echo "== $v"; tail -1 gpurun_out/bench_v.log | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print(round(d['value']), d['ms_per_step'], d['roofline']['achieved'], d['roofline']['share_of_step'], d['clocks']['sm_mhz'])"
done
UNIMM_GELU_TANH=1 timeout 900 $P tests/test_parity_gpu.py -k "fp16" > gpurun_out/p_tanh.log 2>&1; grep -E "passed|failed" gpurun_out/p_tanh.log; grep -E "^\.?\[fp16\]" gpurun_out/p_tanh.log | cut -c1-160
