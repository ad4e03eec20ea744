# ncu evidence for profiles/: (1) every launch with its device time, (2) one full capture of the top kernel.
mkdir -p gpurun_out
set -x
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --precision ${PREC:-fp16}"
$CMD > gpurun_out/prof_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
tail -2 gpurun_out/ncu_launches.log
$CMD > gpurun_out/prof_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:umma_gemm -s 40 -c 4 -o gpurun_out/prof_gemm $CMD > gpurun_out/ncu_gemm.log 2>&1
tail -2 gpurun_out/ncu_gemm.log
$CMD > gpurun_out/prof_plain3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:attn_mma -s 8 -c 3 -o gpurun_out/prof_attn $CMD > gpurun_out/ncu_attn.log 2>&1
tail -2 gpurun_out/ncu_attn.log
ls -la gpurun_out
