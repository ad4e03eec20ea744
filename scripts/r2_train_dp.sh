mkdir -p gpurun_out
N=${NGPU:-2}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --workload train_step --gpus $N --steps 5 --warmup 3 > gpurun_out/r2_train_step_fp16_n$N.json 2> gpurun_out/r2_train_step_fp16_n$N.err
tail -5 gpurun_out/r2_train_step_fp16_n$N.err; cut -c1-300 gpurun_out/r2_train_step_fp16_n$N.json
