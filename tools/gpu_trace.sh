#!/bin/bash
mkdir -p gpurun_out
for v in "$@"; do
  export MIPNERF_B200_LIB=$PWD/mipnerf_pl_b200/libmipnerf_b200.$v.so
  timeout 300 python tools/tc_trace.py pair > gpurun_out/trace_$v.txt 2>&1
  echo "=== $v"; grep -E "^g |^g[0-9]|slot|features|tile period" gpurun_out/trace_$v.txt | head -30
done
