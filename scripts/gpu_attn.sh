mkdir -p gpurun_out
P="python -m pytest -q -s -p no:cacheprovider"
timeout 600 $P tests/test_kernels_gpu.py -k "candidate_attention" > gpurun_out/k_attn.log 2>&1; tail -8 gpurun_out/k_attn.log | cut -c1-300
timeout 900 $P tests/test_parity_gpu.py -k "prefix_shared" > gpurun_out/p_attn.log 2>&1; tail -3 gpurun_out/p_attn.log | cut -c1-300
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_v.log 2>&1; tail -1 gpurun_out/bench_v.log | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('cand/s', round(d['value']), 'ms', round(d['ms_per_step'],2), 'gemm TF', round(d['roofline']['achieved']), d['roofline']['share_of_step'], d['clocks']['sm_mhz'])"
