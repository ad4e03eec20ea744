mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_sweep_parity_gpu.py -q -m gpu -p no:cacheprovider -x -k "in_flight" 2>&1 | grep -E "^E |Error|assert|passed|failed" | head -20
timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-bf16 2>&1 | tail -5 | cut -c1-600
