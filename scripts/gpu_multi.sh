# N-GPU bench line (one rank per GPU, torchrun) next to the reference arm launched the same way; logs into gpurun_out/
mkdir -p gpurun_out
N=${N:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR bench.py --gpus $N --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n$N.log 2>&1
grep '^{' gpurun_out/bench_n$N.log | tail -1 | cut -c1-600
if [ "${REF:-0}" = "1" ]; then
  timeout 300 $TR bench.py --impl reference --gpus $N --steps 5 --warmup 1 > gpurun_out/bench_ref_n$N.log 2>&1
  grep '^{' gpurun_out/bench_ref_n$N.log | tail -1 | cut -c1-300
fi
