# round 2, 11th GPU call: FMA-pipe exp2 for half of pass 2 of the tcgen05 attention: parity (16-bit modes) + A/B timing
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_sweep_parity_gpu.py -q -m gpu -p no:cacheprovider -k "in_flight or truncated" 2>&1 | tail -2
for dbg in 16 0; do
UNIMM_ATTN_DBG=$dbg timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-bf16 > gpurun_out/r2_bench_attn_dbg$dbg.json 2>/dev/null
python -c "
import json; d=json.load(open('gpurun_out/r2_bench_attn_dbg$dbg.json')); print('dbg $dbg', d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['share_of_step'], d['clocks']['sm_mhz'])"
done
UNIMM_ATTN_DBG=16 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-bf16 > gpurun_out/r2_bench_attn_dbg16b.json 2>/dev/null
python -c "
import json; d=json.load(open('gpurun_out/r2_bench_attn_dbg16b.json')); print('dbg 16 again', d['value'], d['ms_per_step'], d['roofline']['share_of_step'])"
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-bf16 > gpurun_out/r2_bench_attn_dbg0b.json 2>/dev/null
python -c "
import json; d=json.load(open('gpurun_out/r2_bench_attn_dbg0b.json')); print('dbg 0 again', d['value'], d['ms_per_step'], d['roofline']['share_of_step'])"
