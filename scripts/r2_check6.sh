# round 2, 6th GPU call: fp32-class mode with the fused LSE epilogue (parity subset + bench), then the ncu evidence of the fp16 step
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_sweep_parity_gpu.py -q -m gpu -p no:cacheprovider -s -k "fp32" > gpurun_out/r2_gpu_fp32.log 2>&1; grep -E "passed|failed" gpurun_out/r2_gpu_fp32.log | tail -2
grep -E "^FAILED|config 1|bench-shape|config 3" gpurun_out/r2_gpu_fp32.log | head
timeout 900 python bench.py --steps 10 --warmup 3 --precision fp32 --no-cpu-baseline > gpurun_out/r2_bench_fp32_tc3.json 2> gpurun_out/r2_bench_fp32_tc3.err; tail -3 gpurun_out/r2_bench_fp32_tc3.err
python -c "
import json; d=json.load(open('gpurun_out/r2_bench_fp32_tc3.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['share_of_step'], d['roofline']['achieved'])"
bash scripts/gpu_profile.sh
PREC=fp32 CMD2=1 bash -c 'CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --precision fp32 --no-bf16"; $CMD > gpurun_out/prof_plain_fp32.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 700 --csv --log-file gpurun_out/launches_fp32.csv $CMD > gpurun_out/ncu_launches_fp32.log 2>&1; tail -1 gpurun_out/ncu_launches_fp32.log | cut -c1-200'
ls -la gpurun_out/*.ncu-rep gpurun_out/launches*.csv
