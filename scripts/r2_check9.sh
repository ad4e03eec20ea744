# round 2, 9th GPU call: async two-slot pipeline: tests, default bench line, whole 2064-image sweep on one GPU
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_lm_head_backward_gpu.py tests/test_sweep_parity_gpu.py tests/test_val_sweep_gpu.py -q -m gpu -p no:cacheprovider -k "lm_head_backward or in_flight or val_sweep or out_of_range" 2>&1 | tail -4
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_v3.json 2> gpurun_out/r2_bench_v3.err; tail -3 gpurun_out/r2_bench_v3.err
python -c "
import json; d=json.load(open('gpurun_out/r2_bench_v3.json')); print(d['value'], d['ms_per_step'], d['e2e'], d['e2e_wall_clock'], d['bf16_mode']['value'], d['bf16_mode']['e2e'], d['clocks'])"
timeout 1200 python bench.py --workload sweep --images 2064 > gpurun_out/r2_sweep2064_n1.json 2> gpurun_out/r2_sweep2064_n1.err; tail -2 gpurun_out/r2_sweep2064_n1.err; python -c "
import json; d=json.load(open('gpurun_out/r2_sweep2064_n1.json')); s=d['sweep']; print('sweep2064 n1', d['value'], s['sweep_seconds'], s['phases_rank0'], s['generation_seconds_rank0'], d['clocks'])"
