# round 2, 4th GPU call: new tests (no -x), pooler GEMM, dense workloads (configs 3/4/5)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider -s > gpurun_out/r2_gpu_all.log 2>&1; grep -E "passed|failed" gpurun_out/r2_gpu_all.log | tail -3
grep -E "^FAILED|bench-shape|truncated|config 4|config 3|config 5" gpurun_out/r2_gpu_all.log | head -60
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
for wl in train_fwd dis_nsp dense_ft; do
  timeout 600 python bench.py --workload $wl --steps 10 --warmup 3 > gpurun_out/r2_bench_$wl.json 2> gpurun_out/r2_bench_$wl.err || tail -5 gpurun_out/r2_bench_$wl.err
  python -c "
import json; d=json.load(open('gpurun_out/r2_bench_$wl.json')); print('$wl', d['value'], d['unit'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['pct_of_bf16_peak']['burst'], d['roofline']['share_of_step'], d['config']['result_of_last_step'])"
done
timeout 600 python bench.py --workload train_fwd --steps 10 --warmup 3 --precision bf16 > gpurun_out/r2_bench_train_fwd_bf16.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/r2_bench_train_fwd_bf16.json')); print('train_fwd bf16', d['value'], d['pct_of_bf16_peak']['burst'])"
