mkdir -p gpurun_out
P="python -m pytest -q -s -p no:cacheprovider -x"
timeout 500 $P tests/test_parity_gpu.py -k "scores_only or prefix_shared" 2>&1 | grep -E "scores-only|passed|failed|FAIL|Error|error" | cut -c1-220
timeout 500 $P tests/test_edges_gpu.py tests/test_fullsize_gpu.py tests/test_val_sweep_gpu.py 2>&1 | grep -E "scores|passed|failed|FAIL|Error" | cut -c1-220
run() {
  env $1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_dd.log 2>&1
  tail -1 gpurun_out/bench_dd.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); r=d['roofline']; print('$1', 'cand/s', round(d['value']), 'e2e', round(d['e2e']['value']), 'ms', round(d['ms_per_step'],2), 'GF/cand', round(d['config']['executed_flops_per_candidate']/1e9,3), r['share_of_step'], d['clocks']['sm_mhz'])" || tail -5 gpurun_out/bench_dd.log
}
run UNIMM_LM_DEDUP=1
run UNIMM_LM_DEDUP=0
run UNIMM_LM_DEDUP=1
