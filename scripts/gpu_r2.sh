# round-1 session-2 check: fused GEMM+LN kernel tests, GEMM microbench, whole suite, bench with and without the fusion
mkdir -p gpurun_out
P="python -m pytest -q -s -p no:cacheprovider"
timeout 600 $P tests/test_kernels_gpu.py -x -k "gemm_ln" > gpurun_out/k_ln.log 2>&1; tail -3 gpurun_out/k_ln.log; grep -E "gemm_ln\[|FAIL|rror" gpurun_out/k_ln.log | head -30
timeout 900 $P tests/test_kernels_gpu.py tests/test_metrics_gpu.py > gpurun_out/k_all.log 2>&1; tail -3 gpurun_out/k_all.log; grep -E "FAIL|Error|error" gpurun_out/k_all.log | head -20
timeout 600 python scripts/gemm_bench.py 64000 ${GB_MODES:-nomodes} > gpurun_out/gemm_bench2.log 2>&1; cat gpurun_out/gemm_bench2.log
timeout 1800 $P tests/test_parity_gpu.py > gpurun_out/p_all.log 2>&1; grep -E "^\.?\[|passed|failed|FAIL|err" gpurun_out/p_all.log | cut -c1-200 | head -40
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ln.log 2>&1; tail -1 gpurun_out/bench_ln.log | cut -c1-300; tail -1 gpurun_out/bench_ln.log | grep -o '"roofline.*' | cut -c1-600
UNIMM_RES16=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_res32.log 2>&1; tail -1 gpurun_out/bench_res32.log | cut -c1-300; tail -1 gpurun_out/bench_res32.log | grep -o "\"roofline.*" | cut -c1-600
