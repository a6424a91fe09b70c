#!/bin/bash
mkdir -p gpurun_out
echo "== TC step vs reference golden (calibration prints)"
timeout 600 python -m pytest tests/test_gpu_training.py -m gpu -q -s -k "vs_reference_autograd_golden" 2>&1 | grep -E "case|passed|failed" | cut -c1-700
echo "== inference bench: default vs rolled epilogue (LDG biases)"
for V in "" rolled; do
  if [ -n "$V" ]; then export MIPNERF_B200_LIB=$PWD/mipnerf_pl_b200/libmipnerf_b200.$V.so; fi
  timeout 600 python bench.py --steps 200 --warmup 20 --no-cpu-baseline --no-frame --no-parity 2>/dev/null | python -c "
import sys, json
for line in sys.stdin:
    if line.startswith('{'):
        d=json.loads(line); print('$V', d['value'], d['ms_per_step'], d.get('kernel_ms_per_step'), d['roofline']['frac'])
"
done
