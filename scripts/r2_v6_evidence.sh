mkdir -p gpurun_out
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_v6_smoke.txt 2>&1; tail -6 gpurun_out/r2_v6_smoke.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2600 --csv --log-file gpurun_out/r2_v6_train_launches.csv python scripts/train_step_once.py > gpurun_out/r2_v6_train_ncu.log 2>&1; tail -2 gpurun_out/r2_v6_train_ncu.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_bwd_dkv_kernel -s 3 -c 1 -o gpurun_out/r2_v6_prof_attn_bwd_dkv python scripts/train_step_once.py > gpurun_out/r2_v6_ncu_full.log 2>&1; tail -2 gpurun_out/r2_v6_ncu_full.log
