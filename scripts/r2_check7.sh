# round 2, 7th GPU call: LM-head backward parity, the whole GPU suite, smoke, DRAM traffic of the dominant GEMM, full default bench line
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_lm_head_backward_gpu.py -q -m gpu -p no:cacheprovider -s 2>&1 | grep -E "passed|failed|\] n=|Error|error" | head -20
timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/r2_gpu_all.log 2>&1; grep -E "passed|failed" gpurun_out/r2_gpu_all.log | tail -2; grep -E "^FAILED" gpurun_out/r2_gpu_all.log | head
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
bash scripts/gpu_traffic.sh
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_v2.json 2> gpurun_out/r2_bench_v2.err; tail -3 gpurun_out/r2_bench_v2.err
python -c "
import json; d=json.load(open('gpurun_out/r2_bench_v2.json')); print(d['value'], d['ms_per_step'], d['e2e'], d['pct_of_bf16_peak'], d['bf16_mode']['value'], d['bf16_mode']['pct_of_bf16_peak_burst'], d['cpu_baseline'])"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_ref.json 2>&1; tail -c 600 gpurun_out/r2_bench_ref.json
