# full GPU test-suite + one bench line (fp16); logs into gpurun_out/
mkdir -p gpurun_out
P="python -m pytest -q -s -p no:cacheprovider"
timeout 900 $P tests/test_kernels_gpu.py > gpurun_out/k_all.log 2>&1; tail -4 gpurun_out/k_all.log; grep -E "FAIL|Error|error" gpurun_out/k_all.log | head -20
timeout 1500 $P tests/test_parity_gpu.py > gpurun_out/p_all.log 2>&1; grep -E "^\[|passed|failed|FAIL" gpurun_out/p_all.log | cut -c1-200
timeout 900 python bench.py --steps ${STEPS:-6} --warmup 3 --no-cpu-baseline --precision ${PREC:-fp16} > gpurun_out/bench.log 2>&1; tail -2 gpurun_out/bench.log | cut -c1-3000
