#!/bin/bash
mkdir -p gpurun_out
export MIPNERF_B200_TC_VARIANT=v3
echo "== v3 tests + stress + bench"
timeout 600 python -m pytest tests/test_gpu_tensorcore.py -m gpu -q -k "v3" 2>&1 | tail -3
timeout 600 python tools/v3_stress.py 4096 100 both 2>&1 | grep -v "^frame" | grep -E "Error|error|ok|differs|timeout|site" | grep -v "iteration" | head
timeout 600 python bench.py --steps 100 --warmup 10 --no-cpu-baseline --no-frame > gpurun_out/r2_bench_v3.json 2> gpurun_out/r2_bench_v3.err; tail -2 gpurun_out/r2_bench_v3.err | cut -c1-300
python -c "
import json; d=json.load(open('gpurun_out/r2_bench_v3.json')); print('v3', d['value'], d['ms_per_step'], d['e2e']['value'], d['kernel_ms'], d['roofline']['launch_ms'], d['roofline']['frac'], d['parity'])"
