#!/bin/bash
# ncu evidence for profiles/: full capture of the fused level kernel + launch list of one bench run
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:mlp_level -s 6 -c 2 -f -o gpurun_out/prof_level python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1; tail -2 gpurun_out/ncu_full.log | cut -c1-200
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-frame > gpurun_out/ncu_list.log 2>&1
ncu -i gpurun_out/prof_level.ncu-rep --page raw --csv > gpurun_out/prof_level_raw.csv 2>/dev/null
python tools/ncu_summarize.py gpurun_out/launches.csv gpurun_out/prof_level_raw.csv gpurun_out/r01
ls -la gpurun_out/prof_level.ncu-rep gpurun_out/launches.csv
