mkdir -p gpurun_out
P="python -m pytest -q -s -p no:cacheprovider"
UNIMM_GEMM_MULTICAST=${MODE2:-2} timeout 300 $P tests/test_kernels_gpu.py -x -k "gemm_umma and 40000 or lm_head" > gpurun_out/k_gemm.log 2>&1; tail -4 gpurun_out/k_gemm.log | cut -c1-300; grep -E "FAIL|rror" gpurun_out/k_gemm.log | head -5
for mc in 2 1; do echo "== pair mode $mc"; UNIMM_GEMM_MULTICAST=$mc timeout 300 python scripts/gemm_bench.py 98176 2>&1 | head -7 | cut -c1-220; done
