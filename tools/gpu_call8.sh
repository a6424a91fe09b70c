#!/bin/bash
mkdir -p gpurun_out
echo "== full gpu suite"; timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -12 | tee gpurun_out/pytest_full.txt
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
echo "== bench default"; timeout 600 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; python -c "
import json; d=json.load(open('gpurun_out/bench_default.json')); print({k:d[k] for k in ('value','ms_per_step','dtype','e2e','clocks','gpu_launches','cpu_baseline')}); print(d['roofline'])"; tail -3 gpurun_out/bench_default.err
echo "== bench reference arm"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 | cut -c1-400
