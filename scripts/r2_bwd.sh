mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_lm_head_backward_gpu.py -q -m gpu -p no:cacheprovider -s -k "whole" 2>&1 | grep -E "passed|failed|whole LM|^E " | head -40
