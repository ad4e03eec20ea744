mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_train_step_gpu.py -x -q -m gpu -p no:cacheprovider -s -k "dropout or tiny or attention_backward" 2>&1 | grep -v "^\s*$" | tail -45 > gpurun_out/r2_dropout.out
tail -45 gpurun_out/r2_dropout.out
