# bf16 mode with mixed operand formats (LayerNorm outputs and the weights reading them in fp16: mix16) with / without the fp32-class
# LM head (lm_hp): the bench-shape parity cases per variant, every bf16 test of the suite on the default, the bf16 bench lines
mkdir -p gpurun_out
run_sweep() { timeout 600 python -m pytest tests/test_sweep_parity_gpu.py -q -m gpu -k "bf16 and bench_step" -s -rxX 2>&1 | grep "bench-shape\|passed\|failed\|xfailed\|xpassed"; }
echo "== mix16 + lm_hp (default)"; run_sweep
echo "== mix16 only"; UNIMM_LM_HP=0 run_sweep
echo "== pure bf16 + lm_hp"; UNIMM_BF16_PURE=1 run_sweep
echo "== pure bf16 (round-2 v7)"; UNIMM_BF16_PURE=1 UNIMM_LM_HP=0 run_sweep
timeout 900 python -m pytest tests -q -m gpu -k "bf16" -s -rxX 2>&1 | grep -v "^$" | tail -80 > gpurun_out/r2_v9_bf16_tests.txt; tail -3 gpurun_out/r2_v9_bf16_tests.txt
timeout 600 python bench.py --precision bf16 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2_v9_bench_bf16.json 2> gpurun_out/r2_v9_bench_bf16.err; cut -c1-200 gpurun_out/r2_v9_bench_bf16.json
UNIMM_LM_HP=0 timeout 600 python bench.py --precision bf16 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2_v9_bench_bf16_lmhp0.json 2> gpurun_out/r2_v9_bench_bf16_lmhp0.err; cut -c1-200 gpurun_out/r2_v9_bench_bf16_lmhp0.json
