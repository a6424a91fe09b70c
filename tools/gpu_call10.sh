#!/bin/bash
mkdir -p gpurun_out
echo "== full gpu suite"; timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -6 | tee gpurun_out/pytest_full.txt
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
echo "== bench default"; timeout 600 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; python -c "
import json; d=json.load(open('gpurun_out/bench_default.json')); print({k:d[k] for k in ('value','ms_per_step','dtype','e2e','clocks','gpu_launches','cpu_baseline','frame')}); print(d['roofline'])"; tail -3 gpurun_out/bench_default.err
echo "== bench fp16"; timeout 300 python bench.py --precision fp16 --no-cpu-baseline > gpurun_out/bench_fp16.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/bench_fp16.json')); print(d['value'], d['roofline']['frac'])"
bash tools/gpu_ncu.sh
