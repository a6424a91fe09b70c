#!/bin/bash
mkdir -p gpurun_out
echo "== trace"; MIPNERF_B200_LIB=$PWD/mipnerf_pl_b200/libmipnerf_b200.trace.so timeout 300 python tools/tc_trace.py pair > gpurun_out/trace_v3.txt 2>&1; grep -E "^slot|IPE|view epi|MMA starts|tile period" gpurun_out/trace_v3.txt | head -20
bash tools/gpu_variants.sh st4 2>&1 | grep -v pytest | tail -3
