# round 2, first GPU call: bf16 A/B (bf16 weights vs fp16-stored weights), the whole GPU suite, fp16 + bf16 bench lines
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv,noheader; nproc; free -g | head -2
SEL="tests/test_parity_gpu.py::test_config1_scores_16bit tests/test_parity_gpu.py::test_prefix_shared_scores_match_reference tests/test_parity_gpu.py::test_scores_only_packing_matches_reference tests/test_parity_gpu.py::test_generative_scores tests/test_parity_gpu.py::test_training_losses"
for w in 0 1; do
  echo "== UNIMM_BF16_WFP16=$w"
  UNIMM_BF16_WFP16=$w timeout 600 python -m pytest $SEL -q -s -m gpu -k bf16 -p no:cacheprovider 2>&1 | grep -E "^\[bf16\]|passed|failed|Error" 
done > gpurun_out/r2_bf16_ab.log 2>&1
cat gpurun_out/r2_bf16_ab.log
timeout 900 python -m pytest tests -q -x -m gpu -p no:cacheprovider > gpurun_out/gpu_all.log 2>&1; tail -5 gpurun_out/gpu_all.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_fp16_base.json 2> gpurun_out/r2_bench_fp16_base.err; tail -c 1500 gpurun_out/r2_bench_fp16_base.json
timeout 600 python bench.py --steps 20 --warmup 5 --precision bf16 --no-cpu-baseline > gpurun_out/r2_bench_bf16_base.json 2> gpurun_out/r2_bench_bf16_base.err; head -c 400 gpurun_out/r2_bench_bf16_base.json
