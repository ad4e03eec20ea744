mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_val_sweep_gpu.py tests/test_lm_head_backward_gpu.py -q -m gpu -p no:cacheprovider 2>&1 | tail -5
