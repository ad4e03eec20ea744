mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_metrics_gpu.py tests/test_train_step_gpu.py -x -q -m gpu -p no:cacheprovider -s -k "ndcg or gradient or tiny" 2>&1 | grep -v "^\s*$" | grep "neuralNDCG\|passed\|failed\|Error\|error\|assert\|tiny\]" | tail -40 > gpurun_out/r2_train4.out
cat gpurun_out/r2_train4.out
