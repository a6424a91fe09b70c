#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_training.py -m gpu -q -k "fused or tracks" 2>&1 | tail -3
timeout 600 python tools/train_bench.py --precision bf16 > gpurun_out/r2_train_bf16_fused6.json 2> gpurun_out/r2_train_bf16_fused6.err; tail -c 1300 gpurun_out/r2_train_bf16_fused6.json; tail -3 gpurun_out/r2_train_bf16_fused6.err
