mkdir -p gpurun_out
set -x
nvidia-smi --query-gpu=name,driver_version,memory.total --format=csv
python -c "import torch;print(torch.__version__, torch.cuda.is_available())"
P="python -m pytest -q -s -p no:cacheprovider"
timeout 600 $P tests/test_kernels_gpu.py -k "gemm_f32 or layernorm or verify" > gpurun_out/k1.log 2>&1; tail -25 gpurun_out/k1.log
timeout 600 $P tests/test_kernels_gpu.py -k "attention" > gpurun_out/k2.log 2>&1; tail -25 gpurun_out/k2.log
timeout 600 $P tests/test_kernels_gpu.py -k "umma" > gpurun_out/k3.log 2>&1; tail -40 gpurun_out/k3.log
timeout 600 $P tests/test_kernels_gpu.py -k "lm_head" > gpurun_out/k4.log 2>&1; tail -25 gpurun_out/k4.log
timeout 1500 $P tests/test_parity_gpu.py -k "fp32 or drop_in or host_buffer" > gpurun_out/p1.log 2>&1; tail -40 gpurun_out/p1.log
timeout 1500 $P tests/test_parity_gpu.py -k "bf16" > gpurun_out/p2.log 2>&1; tail -40 gpurun_out/p2.log
