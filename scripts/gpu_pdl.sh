# programmatic dependent launch: parity tests with PDL on, then A/B bench lines; logs into gpurun_out/
mkdir -p gpurun_out
P="python -m pytest -q -s -p no:cacheprovider -x"
timeout 500 $P tests/test_kernels_gpu.py 2>&1 | tail -2 | cut -c1-200
timeout 500 $P tests/test_parity_gpu.py 2>&1 | grep -E "passed|failed|FAIL|Error" | cut -c1-220
timeout 300 $P tests/test_fullsize_gpu.py 2>&1 | grep -E "packing|packed vs|passed|failed|FAIL|Error" | cut -c1-220
run() {
  env $1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline $2 > gpurun_out/bench_pdl.log 2>&1
  tail -1 gpurun_out/bench_pdl.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1 $2', 'cand/s', round(d['value']), 'e2e', round(d['e2e']['value']), 'ms', round(d['ms_per_step'],2), 'sust', round(d['pct_of_bf16_peak']['sustained'],3), d['clocks']['sm_mhz'], d['gpu_launches'])" || tail -5 gpurun_out/bench_pdl.log
}
run UNIMM_PDL=1 ""
run UNIMM_PDL=0 ""
run UNIMM_PDL=1 ""
run UNIMM_PDL=0 ""
run UNIMM_PDL=1 "--images-per-step 16"
