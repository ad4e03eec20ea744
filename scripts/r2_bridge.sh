mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_reference_callers_gpu.py tests/test_train_step_gpu.py -x -q -m gpu -p no:cacheprovider -s -k "loop_body or config3_training" 2>&1 | grep -v "^\s*$" | tail -40 > gpurun_out/r2_bridge.out
tail -40 gpurun_out/r2_bridge.out
