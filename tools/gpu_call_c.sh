#!/bin/bash
mkdir -p gpurun_out
echo "== fp32 parity + training tests"; timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_training.py -m gpu -q 2>&1 | tail -4
echo "== train bench"; timeout 600 python tools/train_bench.py > gpurun_out/train_bench.json 2> gpurun_out/train_bench.err; tail -2 gpurun_out/train_bench.err; cat gpurun_out/train_bench.json
echo "== fp32 bench"; timeout 600 python bench.py --precision fp32 --steps 5 --warmup 3 --no-cpu-baseline --no-frame > gpurun_out/bench_fp32.json 2> gpurun_out/bench_fp32.err; python -c "
import json; d=json.load(open('gpurun_out/bench_fp32.json')); print(round(d['value']), round(d['ms_per_step'],3), d['kernel_ms'])"
echo "== bf16 bench (e2e with RayStaging)"; timeout 600 python bench.py --no-cpu-baseline --no-frame > gpurun_out/bench_e2e.json 2> gpurun_out/bench_e2e.err; python -c "
import json; d=json.load(open('gpurun_out/bench_e2e.json')); print(round(d['value']), round(d['ms_per_step'],4), d['e2e'], round(d['roofline']['step_frac_of_roofline'],4))"
