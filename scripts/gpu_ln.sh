mkdir -p gpurun_out
P="python -m pytest -q -s -p no:cacheprovider"
timeout 900 $P tests/test_kernels_gpu.py -k "gemm_ln" > gpurun_out/k_ln.log 2>&1; tail -2 gpurun_out/k_ln.log; grep -E "FAIL|rror" gpurun_out/k_ln.log | head -5
for mc in 1 0; do echo "== LN multicast $mc"; UNIMM_LN_MULTICAST=$mc timeout 600 python scripts/gemm_bench.py 98176 2>&1 | tail -8 | grep 98176 | cut -c1-250; done
REPS="1 2" VAR=UNIMM_LN_MULTICAST bash scripts/gpu_ab.sh
