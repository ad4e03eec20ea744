# round-2 v6 single-GPU evidence: the whole GPU suite, the driver's bench line, the training-step lines (fp16 / bf16) with the per-operation table
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider 2>&1 | tail -15 > gpurun_out/r2_v6_gpu_tests.txt
tail -5 gpurun_out/r2_v6_gpu_tests.txt
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_v6_bench_fp16.json 2> gpurun_out/r2_v6_bench_fp16.err; cut -c1-260 gpurun_out/r2_v6_bench_fp16.json
timeout 600 python bench.py --workload train_step --steps 10 --warmup 3 --profile-ops > gpurun_out/r2_v6_train_step_fp16.json 2> gpurun_out/r2_v6_train_step_fp16.err; cut -c1-220 gpurun_out/r2_v6_train_step_fp16.json
timeout 600 python bench.py --workload train_step --steps 10 --warmup 3 --precision bf16 > gpurun_out/r2_v6_train_step_bf16.json 2> gpurun_out/r2_v6_train_step_bf16.err; cut -c1-220 gpurun_out/r2_v6_train_step_bf16.json
