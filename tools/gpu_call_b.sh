#!/bin/bash
mkdir -p gpurun_out
echo "== tc tests (fused prologue)"; timeout 900 python -m pytest tests/test_gpu_tensorcore.py -m gpu -q -x 2>&1 | tail -3
for mode in fused separate fused separate; do
  MIPNERF_B200_TC_PROLOGUE=$mode timeout 600 python bench.py --no-cpu-baseline --no-frame > gpurun_out/bench_$mode.json 2> gpurun_out/bench_$mode.err
  python -c "
import json; d=json.load(open('gpurun_out/bench_$mode.json')); print('$mode', round(d['value']), round(d['ms_per_step'],4), d['kernel_ms'], round(d['roofline']['frac'],4), round(d['roofline']['step_frac_of_roofline'],4), 'e2e', round(d['e2e']['value']))"
done
