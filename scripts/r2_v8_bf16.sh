# bf16 mode with the LM head at the fp32-class precision (lm_hp): parity cases in bf16 + the bf16 bench line, with the A/B (UNIMM_LM_HP=0)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_sweep_parity_gpu.py tests/test_parity_gpu.py -q -m gpu -k "bf16" -s -rxX 2>&1 | grep -v "^$" | tail -60 > gpurun_out/r2_v8_bf16_tests.txt
UNIMM_LM_HP=0 timeout 900 python -m pytest tests/test_sweep_parity_gpu.py -q -m gpu -k "bf16 and bench_step" -s -rxX 2>&1 | grep -v "^$" | tail -20 > gpurun_out/r2_v8_bf16_tests_lmhp0.txt
timeout 600 python bench.py --precision bf16 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2_v8_bench_bf16.json 2> gpurun_out/r2_v8_bench_bf16.err; cut -c1-300 gpurun_out/r2_v8_bench_bf16.json
UNIMM_LM_HP=0 timeout 600 python bench.py --precision bf16 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2_v8_bench_bf16_lmhp0.json 2> gpurun_out/r2_v8_bench_bf16_lmhp0.err; cut -c1-300 gpurun_out/r2_v8_bench_bf16_lmhp0.json
grep -h "bench-shape\|passed\|failed\|XPASS\|XFAIL" gpurun_out/r2_v8_bf16_tests.txt gpurun_out/r2_v8_bf16_tests_lmhp0.txt
