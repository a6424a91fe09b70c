#!/bin/bash
# One GPU call that re-validates the tree: gpu test suite, smoke(), default bench line.
mkdir -p gpurun_out
echo "== gpu suite"; timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -5
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
echo "== bench"; timeout 900 python bench.py > gpurun_out/bench_validate.json 2> gpurun_out/bench_validate.err; python -c "
import json; d=json.load(open('gpurun_out/bench_validate.json')); print({k:d[k] for k in ('value','ms_per_step','e2e','kernel_ms','frame','cpu_baseline','clocks','gpu_launches')}); print(d['roofline'])"; tail -3 gpurun_out/bench_validate.err
