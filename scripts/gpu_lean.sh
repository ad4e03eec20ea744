# scores-only packing + tail pruning: parity tests + A/B bench lines; logs into gpurun_out/
mkdir -p gpurun_out
P="python -m pytest -q -s -p no:cacheprovider"
timeout 600 $P tests/test_parity_gpu.py -k "scores_only or prefix_shared" 2>&1 | grep -E "^\.?\[|passed|failed|FAIL|Error" | cut -c1-220
timeout 600 $P tests/test_fullsize_gpu.py tests/test_val_sweep_gpu.py 2>&1 | grep -E "packing|packed vs|passed|failed|FAIL|Error" | cut -c1-220
for a in 1 0 1; do
  UNIMM_PRUNE_TAIL=$a timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_lean.log 2>&1
  tail -1 gpurun_out/bench_lean.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('prune=$a', 'cand/s', round(d['value']), 'e2e', round(d['e2e']['value']), 'ms', round(d['ms_per_step'],2), 'GF/cand', round(d['config']['executed_flops_per_candidate']/1e9,3), 'rows', d['config']['packed_text_rows_per_step'], 'sust', round(d['pct_of_bf16_peak']['sustained'],3), d['clocks']['sm_mhz'], d['gpu_launches'])" || tail -5 gpurun_out/bench_lean.log
done
