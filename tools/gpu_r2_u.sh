#!/bin/bash
export MIPNERF_B200_TC_VARIANT=v4
echo "== v4 (no trace this run)"
true
echo "== v4 tests + bench"
timeout 600 python -m pytest tests/test_gpu_tensorcore.py -m gpu -q -k "v4" 2>&1 | tail -2
timeout 600 python bench.py --steps 200 --warmup 20 --no-cpu-baseline --no-frame --no-train --no-parity-mode --no-parity 2>/dev/null | python -c "
import sys, json
for line in sys.stdin:
    if line.startswith('{'):
        d=json.loads(line); print('v4', d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['launch_ms'])
"
