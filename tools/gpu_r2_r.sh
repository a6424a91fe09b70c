#!/bin/bash
mkdir -p gpurun_out
echo "== x3 tests"
timeout 600 python -m pytest tests/test_gpu_x3.py -m gpu -q 2>&1 | tail -3
for P in fp16x3 bf16x3 bf16; do
timeout 600 python bench.py --precision $P --steps 200 --warmup 20 --no-cpu-baseline --no-frame --no-train 2>/dev/null | python -c "
import sys, json
for line in sys.stdin:
    if line.startswith('{'):
        d=json.loads(line); print('$P', d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['launch_ms'], (d.get('parity') or {}).get('trained_like'))
"
done
echo "== tensorcore + parity tests"
timeout 900 python -m pytest tests/test_gpu_tensorcore.py tests/test_gpu_parity.py tests/test_gpu_graph.py -m gpu -q 2>&1 | tail -3
