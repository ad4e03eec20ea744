mkdir -p gpurun_out
timeout 200 python -m pytest -q -s -p no:cacheprovider tests/test_val_sweep_gpu.py 2>&1 | tail -2 | cut -c1-250
timeout 300 python -m unimm_b200.val_sweep --images 48 --prefetch 1 2>&1 | tail -1 | cut -c1-600
timeout 300 python -m unimm_b200.val_sweep --images 48 --prefetch 0 2>&1 | tail -1 | cut -c1-600
