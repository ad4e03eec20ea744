mkdir -p gpurun_out
P="python -m pytest -q -s -p no:cacheprovider"
timeout 600 $P tests/test_kernels_gpu.py -k "attention" > gpurun_out/k_attn.log 2>&1; tail -5 gpurun_out/k_attn.log | cut -c1-200; grep -E "FAIL|rror" gpurun_out/k_attn.log | head
timeout 1500 $P tests/test_parity_gpu.py tests/test_fullsize_gpu.py > gpurun_out/p_all.log 2>&1; tail -3 gpurun_out/p_all.log | cut -c1-200; grep -E "FAIL|rror" gpurun_out/p_all.log | head
for m in dense packed; do for v in 1 0; do
  UNIMM_ATTN_UMMA=$v timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --mode $m > gpurun_out/bench_ab.log 2>&1; tail -1 gpurun_out/bench_ab.log | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('$m umma=$v cand/s', round(d['value']), 'ms', round(d['ms_per_step'],2), 'gemm TF', round(d['roofline']['achieved']), d['roofline']['share_of_step'], d['clocks']['sm_mhz'])"
done; done
