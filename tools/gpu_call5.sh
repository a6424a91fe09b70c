#!/bin/bash
mkdir -p gpurun_out
echo "== ncu full (pair, bf16)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:mlp_level -s 6 -c 2 -f -o gpurun_out/prof_pair python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1; tail -3 gpurun_out/ncu_full.log
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_pair.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_list.log 2>&1; tail -2 gpurun_out/ncu_list.log; wc -l gpurun_out/launches_pair.csv
ls -la gpurun_out/*.ncu-rep
