mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_train_step_gpu.py -x -q -m gpu -p no:cacheprovider -s -k "full_config or dense_annotation" 2>&1 | grep -v "^\s*$" | grep "passed\|failed\|Error\|error\|assert\|\[fp16" | tail -40 > gpurun_out/r2_train5.out
cat gpurun_out/r2_train5.out
timeout 600 python bench.py --workload train_step --steps 5 --warmup 3 --profile-ops > gpurun_out/r2_train_step_fp16_c.json 2> gpurun_out/r2_train_step_fp16_c.err
tail -3 gpurun_out/r2_train_step_fp16_c.err; cut -c1-200 gpurun_out/r2_train_step_fp16_c.json
