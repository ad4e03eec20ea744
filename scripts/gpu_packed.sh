mkdir -p gpurun_out
P="python -m pytest -q -s -p no:cacheprovider"
timeout 1500 $P tests/test_parity_gpu.py -k "prefix" > gpurun_out/p_packed.log 2>&1; grep -E "^\[|passed|failed|FAIL|Error|error|assert" gpurun_out/p_packed.log | cut -c1-260 | head -40
