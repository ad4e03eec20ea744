# round 2, 5th GPU call: fp32-class mode with the split (hi | lo, 3-pass mma.sync) attention: parity + bench; N=2 sweep over NCCL is a separate call
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider -s -k "fp32 or float32 or drop_in or reference_callers or val_lm or train_forward or host_buffer or prefix_shared_host or edges or kernels" > gpurun_out/r2_gpu_fp32.log 2>&1; grep -E "passed|failed" gpurun_out/r2_gpu_fp32.log | tail -3
grep -E "^FAILED|\[fp32\]" gpurun_out/r2_gpu_fp32.log | head -50
timeout 900 python bench.py --steps 10 --warmup 3 --precision fp32 --no-cpu-baseline > gpurun_out/r2_bench_fp32_tc2.json 2> gpurun_out/r2_bench_fp32_tc2.err; tail -3 gpurun_out/r2_bench_fp32_tc2.err
python -c "
import json; d=json.load(open('gpurun_out/r2_bench_fp32_tc2.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['share_of_step'], d['roofline']['achieved'])"
timeout 600 python bench.py --workload train_fwd --steps 5 --warmup 2 --precision fp32 > gpurun_out/r2_bench_train_fwd_fp32.json 2> gpurun_out/r2_bench_train_fwd_fp32.err; python -c "
import json; d=json.load(open('gpurun_out/r2_bench_train_fwd_fp32.json')); print('train_fwd fp32', d['value'], d['ms_per_step'], d['roofline']['share_of_step'], d['config']['result_of_last_step'])"
