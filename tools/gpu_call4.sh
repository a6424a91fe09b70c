#!/bin/bash
mkdir -p gpurun_out
echo "== pytest tensorcore (pair+single)"; timeout 600 python -m pytest tests/test_gpu_tensorcore.py -m gpu -q -x 2>&1 | tail -25 | tee gpurun_out/pytest_tc4.txt
echo "== bench bf16 pair"; timeout 300 python bench.py --steps 20 --warmup 5 --precision bf16 --no-cpu-baseline > gpurun_out/bench_bf16_pair.json 2> gpurun_out/bench_bf16_pair.err; python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/bench_bf16_pair.json')); print({k:d[k] for k in ('value','ms_per_step','e2e','clocks','kernel_ms')}); print(d['roofline'])
except Exception as e: print('bench failed', e); print(open('gpurun_out/bench_bf16_pair.err').read()[-2000:])
PY
echo "== bench bf16 single"; MIPNERF_B200_TC_VARIANT=single timeout 300 python bench.py --steps 20 --warmup 5 --precision bf16 --no-cpu-baseline > gpurun_out/bench_bf16_single.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/bench_bf16_single.json')); print(d['value'], d['roofline']['frac'])"
echo "== pytest parity"; timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q 2>&1 | tail -40 | tee gpurun_out/pytest_gpu4.txt | grep -E "Error|passed|failed" 
