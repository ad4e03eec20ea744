mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_lm_head_backward_gpu.py tests/test_train_step_gpu.py -x -q -m gpu -p no:cacheprovider -s 2>&1 | grep -v "^\s*$" | grep "\[\|passed\|failed\|Error\|error\|assert" | tail -70 > gpurun_out/r2_train3.out
tail -50 gpurun_out/r2_train3.out
timeout 600 python bench.py --workload train_step --steps 5 --warmup 3 > gpurun_out/r2_train_step_fp16_b.json 2> gpurun_out/r2_train_step_fp16_b.err
tail -3 gpurun_out/r2_train_step_fp16_b.err; cut -c1-400 gpurun_out/r2_train_step_fp16_b.json
