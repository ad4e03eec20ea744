# round 2, 10th GPU call: compute-sanitizer memcheck over the round-2 kernel paths (tiny config), then the two-pass bench line
mkdir -p gpurun_out
timeout 300 python scripts/sanitize_small.py 2>&1 | tail -12
timeout 1200 compute-sanitizer --tool memcheck --print-limit 20 python scripts/sanitize_small.py > gpurun_out/r2_memcheck.log 2>&1; tail -5 gpurun_out/r2_memcheck.log; grep -c "Invalid\|out of bounds\|misaligned" gpurun_out/r2_memcheck.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_v4.json 2> gpurun_out/r2_bench_v4.err; tail -3 gpurun_out/r2_bench_v4.err
python -c "
import json; d=json.load(open('gpurun_out/r2_bench_v4.json')); print(d['value'], d['ms_per_step'], d['profiled_pass']['ms_per_step'], d['e2e'], d['pct_of_bf16_peak'], d['roofline']['achieved'], d['roofline']['share_of_step'], d['bf16_mode']['value'], d['bf16_mode']['pct_of_bf16_peak_burst'])"
