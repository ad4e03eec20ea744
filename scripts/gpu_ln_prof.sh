mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:umma_gemm_ln -s 2 -c 1 -f -o gpurun_out/prof_ln python scripts/ln_probe.py 768 768 > gpurun_out/ncu_ln.log 2>&1; tail -2 gpurun_out/ncu_ln.log
