mkdir -p gpurun_out
timeout 200 python -m pytest -q -s -p no:cacheprovider tests/test_val_sweep_gpu.py 2>&1 | tail -1 | cut -c1-250
for p in 2 1; do timeout 300 python -m unimm_b200.val_sweep --images 64 --prefetch $p 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('prefetch', d['prefetch'], 'cand/s', round(d['sweep_candidates_per_sec']), 'sec', round(d['sweep_seconds'],3), 'mrr', d['mrr'])"; done
