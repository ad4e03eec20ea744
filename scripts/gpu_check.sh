# full GPU test-suite + one bench line; logs into gpurun_out/
mkdir -p gpurun_out
P="python -m pytest -q -s -p no:cacheprovider"
timeout 900 $P tests/test_kernels_gpu.py > gpurun_out/k_all.log 2>&1; tail -3 gpurun_out/k_all.log; grep -E "FAIL|Error|error" gpurun_out/k_all.log | head -20
timeout 1800 $P tests/test_parity_gpu.py > gpurun_out/p_all.log 2>&1; grep -E "^\.?\[|passed|failed|FAIL" gpurun_out/p_all.log | cut -c1-200
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
timeout 900 python bench.py --steps ${STEPS:-10} --warmup 3 ${BENCH_ARGS:---cpu-sample 25} > gpurun_out/bench.log 2>&1; tail -1 gpurun_out/bench.log | cut -c1-3500
