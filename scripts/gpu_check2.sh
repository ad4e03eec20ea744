mkdir -p gpurun_out
P="python -m pytest -q -s -p no:cacheprovider"
timeout 600 $P tests/test_edges_gpu.py 2>&1 | grep -E "^\.?\[|passed|failed|FAIL|Error" | cut -c1-200
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.log 2>&1; tail -1 gpurun_out/bench.log | cut -c1-300
