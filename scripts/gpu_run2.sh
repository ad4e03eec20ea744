mkdir -p gpurun_out
set -x
P="python -m pytest -q -s -p no:cacheprovider"
timeout 900 $P tests/test_kernels_gpu.py > gpurun_out/k_all.log 2>&1; tail -50 gpurun_out/k_all.log
timeout 1500 $P tests/test_parity_gpu.py > gpurun_out/p_all.log 2>&1; tail -45 gpurun_out/p_all.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -8 gpurun_out/smoke.log
timeout 900 python bench.py --steps 4 --warmup 3 --cpu-sample 25 > gpurun_out/bench_fp16.log 2>&1; tail -5 gpurun_out/bench_fp16.log
timeout 600 python bench.py --steps 4 --warmup 3 --precision bf16 --no-cpu-baseline > gpurun_out/bench_bf16.log 2>&1; tail -3 gpurun_out/bench_bf16.log
nproc; lscpu | grep "Model name"
