#!/bin/bash
mkdir -p gpurun_out
echo "== rng + x3 tests"
timeout 600 python -m pytest tests/test_gpu_rng.py tests/test_gpu_x3.py -m gpu -q 2>&1 | tail -12
echo "== v3 counters"
export MIPNERF_B200_LIB=$PWD/mipnerf_pl_b200/libmipnerf_b200.trace.so
MIPNERF_B200_TC_VARIANT=v3 timeout 300 python tools/v3_counters.py 4096 2>&1 | tail -8
