# round 2, third GPU call: suite again (reference callers fixed, fp32 mode now on tcgen05 split3 GEMMs), fp32 bench A/B
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -x -m gpu -p no:cacheprovider -s > gpurun_out/r2_gpu_all.log 2>&1; grep -E "passed|failed" gpurun_out/r2_gpu_all.log | tail -3
grep -E "bench-shape|truncated|config 4|config 3|config 5|train.forward|visdial_evaluate:|\[fp32\]" gpurun_out/r2_gpu_all.log | head -60
timeout 900 python bench.py --steps 5 --warmup 3 --precision fp32 --no-cpu-baseline > gpurun_out/r2_bench_fp32_tc.json 2> gpurun_out/r2_bench_fp32_tc.err; head -c 600 gpurun_out/r2_bench_fp32_tc.json; tail -3 gpurun_out/r2_bench_fp32_tc.err
python -c "
import json; d=json.load(open('gpurun_out/r2_bench_fp32_tc.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['share_of_step'], d['roofline']['achieved'])"
UNIMM_FP32_SIMT=1 timeout 900 python bench.py --steps 3 --warmup 3 --precision fp32 --no-cpu-baseline > gpurun_out/r2_bench_fp32_simt.json 2> gpurun_out/r2_bench_fp32_simt.err; python -c "
import json; d=json.load(open('gpurun_out/r2_bench_fp32_simt.json')); print(d['value'], d['ms_per_step'], d['roofline']['share_of_step'])"
