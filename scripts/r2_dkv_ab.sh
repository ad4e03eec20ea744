mkdir -p gpurun_out
for v in 0 1 0 1; do
UNIMM_DKV_NW4=$v timeout 600 python bench.py --workload train_step --steps 8 --warmup 3 --profile-ops > gpurun_out/r2_dkv_ab_$v.json 2>/dev/null
python - <<PY
import json
d=json.load(open('gpurun_out/r2_dkv_ab_$v.json'))
print("UNIMM_DKV_NW4=$v", round(d["ms_per_step"],2), "ms; attention_backward", d["ms_per_operation_of_one_step"]["attention_backward"]["ms"], "clock", d["clocks"]["sm_mhz"])
PY
done
