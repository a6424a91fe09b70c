#!/bin/bash
mkdir -p gpurun_out
bash tools/gpu_variants.sh default
echo "== trace"; MIPNERF_B200_LIB=$PWD/mipnerf_pl_b200/libmipnerf_b200.trace.so timeout 300 python tools/tc_trace.py pair > gpurun_out/trace_v3.txt 2>&1; grep -E "^g|slot|tile period" gpurun_out/trace_v3.txt | head -32
