# training step with GELU's pre-activation kept as 16-bit values: the step's GPU tests, then the same-box A/B (UNIMM_PRE16=0 / 1)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_train_step_gpu.py tests/test_reference_callers_gpu.py -q -m gpu -x -s 2>&1 | grep "pre-activation\|passed\|failed\|Error" | tail -12 > gpurun_out/r2_v11_train_tests.txt; cat gpurun_out/r2_v11_train_tests.txt
for v in 1 0 1 0; do
  UNIMM_PRE16=$v timeout 600 python bench.py --workload train_step --steps 10 --warmup 3 > gpurun_out/r2_v11_train_step_pre16_$v.json 2> gpurun_out/r2_v11_train_step_pre16_$v.err
  echo "pre16=$v $(cut -c1-200 gpurun_out/r2_v11_train_step_pre16_$v.json) $(grep -o '"peak_memory_gb": [0-9.]*' gpurun_out/r2_v11_train_step_pre16_$v.json)"
done
