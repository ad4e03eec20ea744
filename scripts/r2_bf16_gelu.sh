mkdir -p gpurun_out
for g in 0 1 0 1; do
UNIMM_GELU_TANH=$g timeout 600 python bench.py --steps 20 --warmup 5 --precision bf16 --no-cpu-baseline > gpurun_out/r2_bf16_gelu$g.json 2>/dev/null
python -c "
import json; d=json.load(open('gpurun_out/r2_bf16_gelu$g.json')); print('gelu_tanh $g', d['value'], d['ms_per_step'], d['pct_of_bf16_peak']['burst'], d['roofline']['share_of_step'])"
done
UNIMM_GELU_TANH=1 timeout 600 python -m pytest tests/test_parity_gpu.py -q -m gpu -p no:cacheprovider -s -k "bf16 and (config1 or prefix_shared or scores_only)" 2>&1 | grep -E "\[bf16\]|passed|failed"
