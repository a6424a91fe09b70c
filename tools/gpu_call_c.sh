#!/bin/bash
mkdir -p gpurun_out
echo "== training tests"; timeout 1200 python -m pytest tests/test_gpu_training.py -m gpu -q 2>&1 | tail -3
echo "== train bench"; timeout 600 python tools/train_bench.py > gpurun_out/train_bench.json 2> gpurun_out/train_bench.err; tail -2 gpurun_out/train_bench.err; python -c "
import json; d=json.load(open('gpurun_out/train_bench.json')); print(d['ms_per_step'], d['rays_per_s'], d['approx_tflops'], d['kernel_ms_per_step'])"
echo "== host profile"; timeout 300 python tools/host_profile.py 2>&1 | head -40
