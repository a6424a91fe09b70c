#!/bin/bash
mkdir -p gpurun_out
echo "== gpu suite"; timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -4
echo "== bench"; timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_b13.json 2> gpurun_out/bench_b13.err; python -c "
import json; d=json.load(open('gpurun_out/bench_b13.json')); print({k:d[k] for k in ('value','ms_per_step','e2e','kernel_ms','kernel_launches','frame')}); print(d['roofline'])"; tail -3 gpurun_out/bench_b13.err
