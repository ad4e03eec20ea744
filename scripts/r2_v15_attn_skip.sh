# candidate-row attention: own-key blocks without any allowed (row, key) pair are skipped (UNIMM_ATTN_DBG=32 = no skipping).
# kernel tests, the microbenchmark both ways (results must be identical), then the same-box A/B of the step
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_fullsize_gpu.py -q -m gpu -k "attention or candidate or fullsize or packed" 2>&1 | tail -3
echo "== skipping (default)"; timeout 300 python scripts/attn_bench.py 1 2 2>&1 | tail -3
echo "== no skipping"; UNIMM_ATTN_DBG=32 timeout 300 python scripts/attn_bench.py 1 2 2>&1 | tail -2
for v in 0 32 0 32; do
  UNIMM_ATTN_DBG=$v timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-bf16 > gpurun_out/r2_v15_bench_fp16_dbg$v.json 2> gpurun_out/r2_v15_bench_fp16_dbg$v.err
  echo "dbg=$v $(cut -c1-150 gpurun_out/r2_v15_bench_fp16_dbg$v.json)"
done
