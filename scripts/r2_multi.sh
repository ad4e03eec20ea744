# round 2, multi-GPU call (N = $1): the driver's default line under torchrun, the strong-scaled sweep, LM-head backward re-check
N=${1:-2}; IMAGES=${2:-512}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 900 $TR bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err; tail -3 gpurun_out/r2_bench_n$N.err
python -c "
import json; d=json.loads(open('gpurun_out/r2_bench_n$N.json').read().strip().splitlines()[-1]); print('N=$N steps:', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'bf16', d['bf16_mode']['value'])"
timeout 900 $TR bench.py --gpus $N --workload sweep --images $IMAGES > gpurun_out/r2_sweep${IMAGES}_n$N.json 2> gpurun_out/r2_sweep_n$N.err; tail -3 gpurun_out/r2_sweep_n$N.err
python -c "
import json; d=json.loads(open('gpurun_out/r2_sweep${IMAGES}_n$N.json').read().strip().splitlines()[-1]); s=d['sweep']; print('N=$N sweep:', d['value'], s['sweep_seconds'], s['phases_rank0'], {k: s[k] for k in ('r@1','mrr','ndcg','ties')}, 'gen_s', s['generation_seconds_rank0'])"
if [ "$N" = "2" ]; then
  timeout 600 python -m pytest tests/test_lm_head_backward_gpu.py tests/test_parity_gpu.py -q -m gpu -p no:cacheprovider -k "lm_head_backward or two_engines" 2>&1 | tail -3
fi
