mkdir -p gpurun_out
rm -f gpurun_out/prof_*.ncu-rep
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --precision fp16 --mode packed --images-per-step 8"
$CMD > gpurun_out/prof_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:attn_ -s 6 -c 8 -o gpurun_out/prof_attn $CMD > gpurun_out/ncu_attn.log 2>&1
tail -1 gpurun_out/ncu_attn.log | cut -c1-200
