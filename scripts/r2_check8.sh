# round 2, 8th GPU call: LM-head backward (with output), fp32-class mode with the LM-head dedup, images-per-step sweep
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_lm_head_backward_gpu.py -q -m gpu -p no:cacheprovider -s 2>&1 | grep -E "passed|failed|\] n=|^E " | head -30
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_sweep_parity_gpu.py tests/test_edges_gpu.py tests/test_fullsize_gpu.py -q -m gpu -p no:cacheprovider -k "fp32 or edges or fullsize" 2>&1 | tail -3
timeout 900 python bench.py --steps 10 --warmup 3 --precision fp32 --no-cpu-baseline > gpurun_out/r2_bench_fp32_tc4.json 2> gpurun_out/r2_bench_fp32_tc4.err; tail -3 gpurun_out/r2_bench_fp32_tc4.err
python -c "
import json; d=json.load(open('gpurun_out/r2_bench_fp32_tc4.json')); print('fp32', d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['share_of_step'])"
for ips in 16 32; do
timeout 900 python bench.py --steps 20 --warmup 5 --images-per-step $ips --no-cpu-baseline --no-bf16 > gpurun_out/r2_bench_ips$ips.json 2> gpurun_out/r2_bench_ips$ips.err; tail -3 gpurun_out/r2_bench_ips$ips.err
python -c "
import json; d=json.load(open('gpurun_out/r2_bench_ips$ips.json')); print('ips $ips', d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks'])"
done
timeout 600 python bench.py --workload sweep --images 256 > gpurun_out/r2_sweep256_n1.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/r2_sweep256_n1.json')); s=d['sweep']; print('sweep256 n1', d['value'], s['sweep_seconds'], s['phases_rank0'])"
