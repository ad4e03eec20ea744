# shared B_0 row: parity tests, then same-box A/B bench lines
mkdir -p gpurun_out
P="python -m pytest -q -s -p no:cacheprovider -x"
timeout 500 $P tests/test_parity_gpu.py -k "scores_only or prefix_shared" 2>&1 | grep -E "scores-only|passed|failed|FAIL|Error" | cut -c1-220
timeout 500 $P tests/test_edges_gpu.py tests/test_fullsize_gpu.py tests/test_val_sweep_gpu.py 2>&1 | grep -E "scores|empty|passed|failed|FAIL|Error" | cut -c1-220
run() {
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline $1 > gpurun_out/bench_b0.log 2>&1
  tail -1 gpurun_out/bench_b0.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); r=d['roofline']; print('$1', 'cand/s', round(d['value']), 'e2e', round(d['e2e']['value']), 'ms', round(d['ms_per_step'],2), 'GF/cand', round(d['config']['executed_flops_per_candidate']/1e9,3), 'rows', d['config']['packed_text_rows_per_step'], 'sust', round(d['pct_of_bf16_peak']['sustained'],3), d['clocks']['sm_mhz'])" || tail -5 gpurun_out/bench_b0.log
}
run ""
run "--own-b0"
run ""
