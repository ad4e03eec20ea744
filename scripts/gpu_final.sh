# round-end rehearsal: full GPU suite, smoke, both bench arms, ncu evidence; logs into gpurun_out/
mkdir -p gpurun_out
P="python -m pytest -q -p no:cacheprovider"
timeout 900 $P tests -m gpu -x > gpurun_out/gpu_all.log 2>&1; tail -3 gpurun_out/gpu_all.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -4 gpurun_out/smoke.log
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_ref.log 2>&1; tail -1 gpurun_out/bench_ref.log | cut -c1-200
SKIP=400 bash scripts/gpu_profile.sh
SKIP=64 COUNT=64 bash scripts/gpu_traffic.sh > gpurun_out/traffic_run.log 2>&1; tail -2 gpurun_out/traffic_run.log | cut -c1-300
if [ "${EXTRA:-0}" = "1" ]; then
  timeout 600 python bench.py --steps 10 --warmup 3 --precision bf16 --no-cpu-baseline > gpurun_out/bench_bf16.log 2>&1; tail -1 gpurun_out/bench_bf16.log | cut -c1-200
  timeout 600 python bench.py --steps 10 --warmup 3 --mode dense --no-cpu-baseline > gpurun_out/bench_dense.log 2>&1; tail -1 gpurun_out/bench_dense.log | cut -c1-200
  timeout 600 python bench.py --steps 10 --warmup 3 --nsp-rows --no-cpu-baseline > gpurun_out/bench_nsp.log 2>&1; tail -1 gpurun_out/bench_nsp.log | cut -c1-200
fi
nproc; lscpu | grep "Model name"
