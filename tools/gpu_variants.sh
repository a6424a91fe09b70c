#!/bin/bash
# usage: tools/gpu_variants.sh <variant> ...   ("default" = the in-tree library)
mkdir -p gpurun_out
echo "== pytest tensorcore"; timeout 600 python -m pytest tests/test_gpu_tensorcore.py -m gpu -q -x 2>&1 | tail -4
for v in "$@"; do
  if [ "$v" = "default" ]; then unset MIPNERF_B200_LIB; else export MIPNERF_B200_LIB=$PWD/mipnerf_pl_b200/libmipnerf_b200.$v.so; fi
  for variant in shared pair; do
    MIPNERF_B200_TC_VARIANT=$variant timeout 300 python bench.py --steps 20 --warmup 5 --precision bf16 --no-cpu-baseline > gpurun_out/bench_$v_$variant.json 2> gpurun_out/bench_$v_$variant.err
    python - "$v" "$variant" gpurun_out/bench_$v_$variant.json <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[3])); r=d['roofline']
    print(f"{sys.argv[1]:10s} {sys.argv[2]:6s} rays/s={d['value']:.4g} launch_ms={r['launch_ms']:.4f} frac={r['frac']:.4f} e2e={d['e2e']['value']:.4g}")
except Exception as e:
    print(sys.argv[1], sys.argv[2], 'FAILED', e)
PY
  done
done
