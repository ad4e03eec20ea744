# round-2 v18 (final) single-GPU validation: the whole GPU suite, smoke(), the driver's default bench line, the reference arm, the bf16 / training-step lines
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider -rxXs 2>&1 | tail -15 > gpurun_out/r2_v18_gpu_tests.txt
tail -6 gpurun_out/r2_v18_gpu_tests.txt
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_v18_smoke.txt 2>&1; tail -4 gpurun_out/r2_v18_smoke.txt
timeout 600 python bench.py > gpurun_out/r2_v18_bench_default.json 2> gpurun_out/r2_v18_bench_default.err; cut -c1-260 gpurun_out/r2_v18_bench_default.json
timeout 600 python bench.py --workload train_step --steps 10 --warmup 3 --profile-ops > gpurun_out/r2_v18_train_step_fp16.json 2> gpurun_out/r2_v18_train_step_fp16.err; cut -c1-220 gpurun_out/r2_v18_train_step_fp16.json
timeout 600 python bench.py --workload train_step --steps 10 --warmup 3 --precision bf16 > gpurun_out/r2_v18_train_step_bf16.json 2> gpurun_out/r2_v18_train_step_bf16.err; cut -c1-220 gpurun_out/r2_v18_train_step_bf16.json
