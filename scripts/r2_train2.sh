mkdir -p gpurun_out
timeout 900 python bench.py --workload train_step --steps 5 --warmup 3 > gpurun_out/r2_train_step_fp16.json 2> gpurun_out/r2_train_step_fp16.err
tail -3 gpurun_out/r2_train_step_fp16.err; cat gpurun_out/r2_train_step_fp16.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r2_train_launches.csv python bench.py --workload train_step --steps 1 --warmup 1 > gpurun_out/r2_train_ncu.log 2>&1
tail -2 gpurun_out/r2_train_ncu.log
