#!/bin/bash
mkdir -p gpurun_out
echo "== training tests"; timeout 600 python -m pytest tests/test_gpu_training.py -m gpu -q -x 2>&1 | tail -4
timeout 300 python tools/train_bench.py --precision bf16 > gpurun_out/train_bench_bf16.json 2> gpurun_out/train_bench_bf16.err
python -c "
import json; d=json.load(open('gpurun_out/train_bench_bf16.json')); print('bf16', round(d['ms_per_step'],2), round(d['rays_per_s']), {k:v for k,v in d['kernel_ms_per_step'].items() if v>1})"
