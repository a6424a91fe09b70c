#!/bin/bash
mkdir -p gpurun_out
echo "== pytest tensorcore (shared first)"; timeout 600 python -m pytest tests/test_gpu_tensorcore.py -m gpu -q -x 2>&1 | tail -8
for variant in shared pair; do
  MIPNERF_B200_TC_VARIANT=$variant timeout 300 python bench.py --steps 40 --warmup 10 --precision bf16 --no-cpu-baseline > gpurun_out/bench_$variant.json 2> gpurun_out/bench_$variant.err
  python - $variant <<'PY'
import json,sys
try:
    d=json.load(open(f'gpurun_out/bench_{sys.argv[1]}.json')); r=d['roofline']
    print(f"{sys.argv[1]:8s} rays/s={d['value']:.4g} launch_ms={r['launch_ms']:.4f} frac={r['frac']:.4f} step_frac={r['step_frac_of_roofline']:.4f} e2e={d['e2e']['value']:.4g} frame={d['frame']['ms']:.1f}")
except Exception as e:
    print(sys.argv[1], 'FAILED', e); print(open(f'gpurun_out/bench_{sys.argv[1]}.err').read()[-1500:])
PY
done
echo "== trace shared"; MIPNERF_B200_LIB=$PWD/mipnerf_pl_b200/libmipnerf_b200.trace.so timeout 300 python tools/tc_trace.py shared > gpurun_out/trace_shared.txt 2>&1; grep -E "^g|slot|tile period" gpurun_out/trace_shared.txt | head -30
