#!/bin/bash
# first GPU call: building-block diagnostics, parity tests, short bench
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
echo "== tc_diag" ; timeout 300 python tools/tc_diag.py > gpurun_out/tc_diag.txt 2>&1; tail -40 gpurun_out/tc_diag.txt
echo "== pytest gpu"; timeout 900 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_tc.py 2>&1 | tail -30 | tee gpurun_out/pytest_gpu.txt
echo "== pytest tc"; timeout 300 python -m pytest tests/test_gpu_tc.py -m gpu -q 2>&1 | tail -15 | tee gpurun_out/pytest_tc.txt
echo "== bench"; timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_fp32.json 2> gpurun_out/bench_fp32.err; tail -3 gpurun_out/bench_fp32.json; tail -5 gpurun_out/bench_fp32.err
