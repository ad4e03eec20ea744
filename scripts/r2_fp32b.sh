mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_sweep_parity_gpu.py tests/test_edges_gpu.py tests/test_reference_callers_gpu.py -q -m gpu -p no:cacheprovider -k "fp32 or edges or reference_callers or drop_in or val_lm or train_forward" 2>&1 | tail -3
timeout 900 python bench.py --steps 10 --warmup 3 --precision fp32 --no-cpu-baseline > gpurun_out/r2_bench_fp32_v6.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/r2_bench_fp32_v6.json')); print('fp32', d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['share_of_step'])"
timeout 600 python bench.py --workload train_fwd --steps 5 --warmup 2 --precision fp32 > gpurun_out/r2_bench_train_fwd_fp32_v6.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/r2_bench_train_fwd_fp32_v6.json')); print('train_fwd fp32', d['value'], d['ms_per_step'], d['roofline']['share_of_step'])"
