mkdir -p gpurun_out
P="python -m pytest -q -s -p no:cacheprovider"
timeout 600 $P tests/test_kernels_gpu.py -k "candidate_attention or text_to_image" > gpurun_out/k_attn.log 2>&1; tail -12 gpurun_out/k_attn.log | cut -c1-300
timeout 900 $P tests/test_parity_gpu.py -k "prefix_shared" > gpurun_out/p_attn.log 2>&1; tail -3 gpurun_out/p_attn.log | cut -c1-300
VAR=UNIMM_ATTN_UMMA bash scripts/gpu_ab.sh
