#!/bin/bash
mkdir -p gpurun_out
echo "== linear_tc + tc training tests"; timeout 600 python -m pytest tests/test_gpu_training.py -m gpu -q -x -s -k "linear_tc or tensor_core" 2>&1 | grep -v "^$" | tail -25
echo "== train bench bf16"; timeout 300 python tools/train_bench.py --precision bf16 > gpurun_out/train_bench_bf16.json 2> gpurun_out/train_bench_bf16.err; tail -3 gpurun_out/train_bench_bf16.err; cat gpurun_out/train_bench_bf16.json
