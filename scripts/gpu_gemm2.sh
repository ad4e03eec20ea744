mkdir -p gpurun_out
P="python -m pytest -q -s -p no:cacheprovider"
timeout 900 $P tests/test_kernels_gpu.py -k "gemm or lm_head" > gpurun_out/k_gemm.log 2>&1; tail -3 gpurun_out/k_gemm.log; grep -E "FAIL|rror" gpurun_out/k_gemm.log | head
for mc in 1 0; do echo "== multicast $mc"; UNIMM_GEMM_MULTICAST=$mc timeout 600 python scripts/gemm_bench.py 98176 2>&1 | head -8 | cut -c1-220; done
