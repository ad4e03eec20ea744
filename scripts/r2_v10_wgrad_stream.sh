# training step with the wgrad GEMMs on a second stream: GPU tests of the step, then the same-box A/B of bench --workload train_step
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_train_step_gpu.py tests/test_reference_callers_gpu.py -q -m gpu -x 2>&1 | tail -5 > gpurun_out/r2_v10_train_tests.txt; cat gpurun_out/r2_v10_train_tests.txt
for v in 1 0 1 0; do
  UNIMM_WGRAD_STREAM=$v timeout 600 python bench.py --workload train_step --steps 10 --warmup 3 > gpurun_out/r2_v10_train_step_ws$v.json 2> gpurun_out/r2_v10_train_step_ws$v.err
  echo "wgrad_stream=$v $(cut -c1-200 gpurun_out/r2_v10_train_step_ws$v.json)"
done
