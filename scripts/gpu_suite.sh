mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -x -m gpu -p no:cacheprovider > gpurun_out/gpu_all.log 2>&1; tail -3 gpurun_out/gpu_all.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
