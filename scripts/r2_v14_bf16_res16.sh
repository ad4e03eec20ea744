# bf16 mode on the fp16 mode's fastest LayerNorm-GEMM path: 16-bit (fp16) residual stream added on the tensor core, cta_group::2 pairs.
# every bf16 test of the suite, then the same-box A/B against the fp32 residual stream (UNIMM_RES16=0)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -k "bf16" -s -rxX 2>&1 | grep "\[bf16\|passed\|failed\|Error\|error" | cut -c1-230 | tail -70 > gpurun_out/r2_v14_bf16_tests.txt; grep "bench-shape\|config 1\|range\|passed\|failed\|switch" gpurun_out/r2_v14_bf16_tests.txt
for v in 1 0 1 0; do
  UNIMM_RES16=$v timeout 600 python bench.py --precision bf16 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2_v14_bench_bf16_res16_$v.json 2> gpurun_out/r2_v14_bench_bf16_res16_$v.err
  echo "res16=$v $(cut -c1-150 gpurun_out/r2_v14_bench_bf16_res16_$v.json)"
done
