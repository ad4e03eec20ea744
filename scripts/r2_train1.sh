mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_train_step_gpu.py -x -q -m gpu -p no:cacheprovider -s -k "train_step" 2>&1 | tail -80 > gpurun_out/r2_train1.out
tail -60 gpurun_out/r2_train1.out
