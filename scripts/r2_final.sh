# round 2, final single-GPU call: whole GPU suite, smoke, the driver's two arms, fp32 / dense workload lines with the final build
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/r2_gpu_final.log 2>&1; grep -E "passed|failed" gpurun_out/r2_gpu_final.log | tail -2; grep -E "^FAILED" gpurun_out/r2_gpu_final.log | head
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_v5.json 2> gpurun_out/r2_bench_v5.err; tail -3 gpurun_out/r2_bench_v5.err
python -c "
import json; d=json.load(open('gpurun_out/r2_bench_v5.json')); print(d['value'], d['ms_per_step'], d['profiled_pass']['ms_per_step'], d['e2e'], d['e2e_wall_clock'], d['pct_of_bf16_peak'], d['roofline']['achieved'], d['roofline']['frac'], d['roofline']['share_of_step'], d['bf16_mode']['value'], d['bf16_mode']['pct_of_bf16_peak_burst'], d['cpu_baseline'], d['clocks'])"
timeout 900 python bench.py --steps 10 --warmup 3 --precision fp32 --no-cpu-baseline > gpurun_out/r2_bench_fp32_v5.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/r2_bench_fp32_v5.json')); print('fp32', d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['share_of_step'])"
timeout 900 python bench.py --steps 10 --warmup 3 --precision bf16 --no-cpu-baseline > gpurun_out/r2_bench_bf16_v5.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/r2_bench_bf16_v5.json')); print('bf16', d['value'], d['ms_per_step'], d['e2e']['value'], d['pct_of_bf16_peak'])"
