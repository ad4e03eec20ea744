mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_train_step_gpu.py tests/test_lm_head_backward_gpu.py -x -q -m gpu -p no:cacheprovider -s 2>&1 | grep -v "^\s*$" | grep "passed\|failed\|Error\|error\|assert\|109 grad\|509 grad\|norms vs" | tail -40 > gpurun_out/r2_train6.out
cat gpurun_out/r2_train6.out
timeout 600 python bench.py --workload train_step --steps 5 --warmup 3 --profile-ops > gpurun_out/r2_train_step_fp16_d.json 2> gpurun_out/r2_train_step_fp16_d.err
tail -3 gpurun_out/r2_train_step_fp16_d.err; cut -c1-200 gpurun_out/r2_train_step_fp16_d.json
