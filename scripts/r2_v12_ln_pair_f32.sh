# LayerNorm-fused GEMM with the fp32 residual on cta_group::2 pairs (UNIMM_LN_PAIR_F32=1): kernel tests, bf16 parity at bench shape, same-box A/B
mkdir -p gpurun_out
UNIMM_LN_PAIR_F32=1 timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "gemm_ln" -s 2>&1 | grep "gemm_ln\|passed\|failed" | tail -20
UNIMM_LN_PAIR_F32=1 timeout 600 python -m pytest tests/test_sweep_parity_gpu.py -q -m gpu -k "bf16 and bench_step" -s 2>&1 | grep "bench-shape\|passed\|failed"
for v in 1 0 1 0; do
  UNIMM_LN_PAIR_F32=$v timeout 600 python bench.py --precision bf16 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2_v12_bench_bf16_pair$v.json 2> gpurun_out/r2_v12_bench_bf16_pair$v.err
  echo "pair_f32=$v $(cut -c1-150 gpurun_out/r2_v12_bench_bf16_pair$v.json)"
done
UNIMM_BF16_PURE=1 timeout 600 python bench.py --precision bf16 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2_v12_bench_bf16_pure.json 2> gpurun_out/r2_v12_bench_bf16_pure.err; echo "pure bf16 $(cut -c1-150 gpurun_out/r2_v12_bench_bf16_pure.json)"
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-bf16 > gpurun_out/r2_v12_bench_fp16.json 2> gpurun_out/r2_v12_bench_fp16.err; echo "fp16 $(cut -c1-150 gpurun_out/r2_v12_bench_fp16.json)"
