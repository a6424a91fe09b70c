#!/bin/bash
mkdir -p gpurun_out
export MIPNERF_B200_TC_VARIANT=v3
echo "== v3 tensorcore tests"
timeout 600 python -m pytest tests/test_gpu_tensorcore.py -m gpu -q -x -k "v3" 2>&1 | tail -5
echo "== bench v3"
timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-frame --no-parity > gpurun_out/r2_bench_v3.json 2> gpurun_out/r2_bench_v3.err; tail -3 gpurun_out/r2_bench_v3.err
python -c "
import json; d=json.load(open('gpurun_out/r2_bench_v3.json')); print('v3', d['value'], d['ms_per_step'], d['e2e']['value'], d['kernel_ms'], d['roofline']['launch_ms'], d['roofline']['frac'])"
if [ -f mipnerf_pl_b200/libmipnerf_b200.trace.so ]; then
export MIPNERF_B200_LIB=$PWD/mipnerf_pl_b200/libmipnerf_b200.trace.so
timeout 300 python tools/v3_counters.py 4096 2>&1 | tail -6
fi
