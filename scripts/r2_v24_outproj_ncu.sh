# ncu full capture of ONE training-forward output projection (t0: M = 61440, N = 768, K = 768, dropout + residual epilogue, fp32 out)
mkdir -p gpurun_out
timeout 400 ncu --set full --clock-control none --import-source on -k regex:umma_gemm_kernel -s 2 -c 3 -o gpurun_out/r2_v24_prof_train_fwd_gemms python scripts/train_step_once.py > gpurun_out/r2_v24_ncu.log 2>&1; tail -2 gpurun_out/r2_v24_ncu.log | cut -c1-200
