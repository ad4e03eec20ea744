mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-bf16"
$CMD > gpurun_out/prof_plain.log 2>&1 &&
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:umma_gemm_kernel -s ${SKIP:-62} -c ${COUNT:-62} --csv --log-file gpurun_out/traffic.csv $CMD > gpurun_out/ncu_traffic.log 2>&1
tail -1 gpurun_out/ncu_traffic.log | cut -c1-120
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-bf16 > gpurun_out/bench_v.log 2>&1; tail -1 gpurun_out/bench_v.log | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print(d['roofline'])"
