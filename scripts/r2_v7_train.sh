mkdir -p gpurun_out
timeout 600 python bench.py --workload train_step --steps 10 --warmup 3 --profile-ops > gpurun_out/r2_v7_train_step_fp16.json 2> gpurun_out/r2_v7_train_step_fp16.err; cut -c1-220 gpurun_out/r2_v7_train_step_fp16.json
timeout 600 python bench.py --workload train_step --steps 10 --warmup 3 --dropout 0 > gpurun_out/r2_v7_train_step_fp16_nodrop.json 2> gpurun_out/r2_v7_train_step_fp16_nodrop.err; cut -c1-220 gpurun_out/r2_v7_train_step_fp16_nodrop.json
timeout 600 python bench.py --workload train_step --steps 10 --warmup 3 --precision bf16 > gpurun_out/r2_v7_train_step_bf16.json 2> gpurun_out/r2_v7_train_step_bf16.err; cut -c1-220 gpurun_out/r2_v7_train_step_bf16.json
