mkdir -p gpurun_out
timeout 600 python scripts/gemm_bench.py 64000 modes > gpurun_out/gemm_bench.log 2>&1; cat gpurun_out/gemm_bench.log
timeout 300 python -m pytest -q -p no:cacheprovider tests/test_kernels_gpu.py -k "umma or lm_head" 2>&1 | tail -3
