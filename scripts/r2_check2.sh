# round 2, second GPU call: whole GPU suite with the north-star tolerances, new bench line (flat-layout e2e, nested bf16), a short sweep
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -x -m gpu -p no:cacheprovider -s > gpurun_out/r2_gpu_all.log 2>&1; tail -4 gpurun_out/r2_gpu_all.log
grep -E "^\[|\] " gpurun_out/r2_gpu_all.log | grep -E "bf16\]|truncated|bench-shape|config 4" | head -40
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_v1.json 2> gpurun_out/r2_bench_v1.err; tail -c 2500 gpurun_out/r2_bench_v1.json; tail -5 gpurun_out/r2_bench_v1.err
timeout 600 python bench.py --workload sweep --images 128 > gpurun_out/r2_sweep128.json 2> gpurun_out/r2_sweep128.err; cat gpurun_out/r2_sweep128.json | head -c 3000; tail -5 gpurun_out/r2_sweep128.err
