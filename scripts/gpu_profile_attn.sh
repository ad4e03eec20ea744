# launch list + full capture of the attention kernels only
mkdir -p gpurun_out
rm -f gpurun_out/prof_*.ncu-rep gpurun_out/launches.csv
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --precision ${PREC:-fp16} --mode ${MODE:-packed} --images-per-step ${IPS:-8}"
$CMD > gpurun_out/prof_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s ${SKIP:-400} -c 700 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
tail -1 gpurun_out/ncu_launches.log | cut -c1-200
$CMD > gpurun_out/prof_plain3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:attn_ -s ${ASKIP:-4} -c ${ACOUNT:-8} -o gpurun_out/prof_attn $CMD > gpurun_out/ncu_attn.log 2>&1
tail -1 gpurun_out/ncu_attn.log | cut -c1-200
