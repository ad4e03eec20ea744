# narrow resident identity for the 16-bit residual of the LayerNorm-fused GEMM: kernel + parity tests, then same-box A/B vs the previous build
mkdir -p gpurun_out
P="python -m pytest -q -s -p no:cacheprovider -x"
timeout 300 $P tests/test_kernels_gpu.py -k "gemm_ln" 2>&1 | tail -3 | cut -c1-200
timeout 500 $P tests/test_parity_gpu.py -k "fp16" 2>&1 | grep -E "passed|failed|FAIL|Error" | cut -c1-220
run() {
  env $1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_id.log 2>&1
  tail -1 gpurun_out/bench_id.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); r=d['roofline']; print('$1', 'cand/s', round(d['value']), 'e2e', round(d['e2e']['value']), 'ms', round(d['ms_per_step'],2), 'gemm_ln TF', round(list(r['other_tensor_kernels_tflops'].values())[0]), r['share_of_step'], d['clocks']['sm_mhz'])" || tail -5 gpurun_out/bench_id.log
}
run A=new
run UNIMM_LIB_PATH=$PWD/unimm_b200/lib/libunimm_b200_prev.so
run A=new
run UNIMM_LIB_PATH=$PWD/unimm_b200/lib/libunimm_b200_prev.so
