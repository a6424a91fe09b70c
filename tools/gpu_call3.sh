#!/bin/bash
mkdir -p gpurun_out
echo "== pytest tensorcore"; timeout 600 python -m pytest tests/test_gpu_tensorcore.py -m gpu -q -x -s 2>&1 | tail -40 | tee gpurun_out/pytest_tc3.txt
echo "== pytest parity"; timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q 2>&1 | tail -15 | tee gpurun_out/pytest_gpu3.txt
echo "== bench bf16"; timeout 600 python bench.py --steps 20 --warmup 5 --precision bf16 > gpurun_out/bench_bf16.json 2> gpurun_out/bench_bf16.err; tail -c 2500 gpurun_out/bench_bf16.json; tail -5 gpurun_out/bench_bf16.err
echo "== bench fp16"; timeout 600 python bench.py --steps 20 --warmup 5 --precision fp16 --no-cpu-baseline > gpurun_out/bench_fp16.json 2> gpurun_out/bench_fp16.err; tail -c 1200 gpurun_out/bench_fp16.json; tail -5 gpurun_out/bench_fp16.err
