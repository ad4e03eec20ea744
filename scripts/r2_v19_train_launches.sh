# ncu launch list of ONE training step of the final version (scripts/train_step_once.py), after the same command ran clean without ncu
mkdir -p gpurun_out
timeout 300 python scripts/train_step_once.py > gpurun_out/r2_v19_train_plain.log 2>&1 && tail -1 gpurun_out/r2_v19_train_plain.log | cut -c1-200 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2600 --csv --log-file gpurun_out/r2_v19_train_launches.csv python scripts/train_step_once.py > gpurun_out/r2_v19_train_ncu.log 2>&1; tail -2 gpurun_out/r2_v19_train_ncu.log | cut -c1-200
