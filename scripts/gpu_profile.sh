# ncu evidence for profiles/: (1) every launch with its device time, (2) full captures of the top kernels.
mkdir -p gpurun_out
rm -f gpurun_out/prof_*.ncu-rep gpurun_out/launches.csv
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --precision ${PREC:-fp16} --mode ${MODE:-packed} --images-per-step ${IPS:-8} --no-bf16 ${EXTRA:-}"
$CMD > gpurun_out/prof_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s ${SKIP:-400} -c 700 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
tail -1 gpurun_out/ncu_launches.log | cut -c1-200
$CMD > gpurun_out/prof_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:umma_gemm -s ${GSKIP:-60} -c ${GCOUNT:-14} -o gpurun_out/prof_gemm $CMD > gpurun_out/ncu_gemm.log 2>&1
tail -1 gpurun_out/ncu_gemm.log | cut -c1-200
$CMD > gpurun_out/prof_plain3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:attn_ -s ${ASKIP:-4} -c ${ACOUNT:-8} -o gpurun_out/prof_attn $CMD > gpurun_out/ncu_attn.log 2>&1
tail -1 gpurun_out/ncu_attn.log | cut -c1-200
