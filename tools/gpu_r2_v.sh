#!/bin/bash
mkdir -p gpurun_out
echo "== ncu of the final training kernels (second step: 2 level kernels + the first backward kernels)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"mlp_level_kernel|linear_t16_kernel|wgrad_mn_kernel" -s 40 -c 14 -f -o gpurun_out/r2_prof_train_final python tools/ncu_train_target.py bf16 4096 2>&1 | tail -1
echo "== launch list of one final training step"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_train_launches_final.csv python tools/ncu_train_target.py bf16 4096 > /dev/null 2>&1; wc -l gpurun_out/r2_train_launches_final.csv
