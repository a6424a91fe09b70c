#!/bin/bash
mkdir -p gpurun_out
echo "== ncu x3"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mlp_level -s 4 -c 1 -f -o gpurun_out/r2_prof_fp16x3 python tools/ncu_target.py fp16x3 4096 4 2>&1 | tail -2
echo "== ncu bf16 v1"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mlp_level -s 4 -c 2 -f -o gpurun_out/r2_prof_bf16 python tools/ncu_target.py bf16 4096 4 2>&1 | tail -2
ls -la gpurun_out/*.ncu-rep
