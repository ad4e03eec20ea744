# training forward of FFN-1 through the fragment-ordered (smem-free) epilogue with both 16-bit outputs: the step's GPU tests, same-box A/B
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_train_step_gpu.py tests/test_reference_callers_gpu.py -q -m gpu -x -s 2>&1 | grep "pre-activation\|passed\|failed\|Error\|assert" | tail -12
for v in 1 0 1 0; do
  UNIMM_PRE16_PERM=$v timeout 600 python bench.py --workload train_step --steps 10 --warmup 3 --profile-ops > gpurun_out/r2_v17_train_step_perm$v.json 2> gpurun_out/r2_v17_train_step_perm$v.err
  echo "perm=$v $(cut -c1-200 gpurun_out/r2_v17_train_step_perm$v.json)"
done
python - <<'PY'
import json
for v in (1, 0):
    d = json.loads(open(f"gpurun_out/r2_v17_train_step_perm{v}.json").read().strip().split("\n")[-1])
    t = d["ms_per_operation_of_one_step"]
    print(v, {k: x for k, x in t.items() if k.startswith("linear M=61440 N=3072")})
PY
