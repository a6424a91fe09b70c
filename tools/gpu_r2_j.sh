#!/bin/bash
# wgrad (MN-major) bring-up + the x3 epilogue change
mkdir -p gpurun_out
echo "== wgrad unit test (gen 1 and 2)"
timeout 600 python -m pytest tests/test_gpu_training.py -m gpu -q -k "wgrad_tc_matches" 2>&1 | tail -8
echo "== same with swapped LBO/SBO (diagnostic)"
MIPNERF_B200_WGRAD_SWAP=1 timeout 600 python -m pytest tests/test_gpu_training.py -m gpu -q -k "wgrad_tc_matches and 2-" 2>&1 | tail -4
echo "== training tests"
timeout 900 python -m pytest tests/test_gpu_training.py -m gpu -q 2>&1 | tail -5
echo "== train bench bf16"
timeout 600 python tools/train_bench.py --precision bf16 > gpurun_out/r2_train_bf16.json 2> gpurun_out/r2_train_bf16.err; tail -c 1200 gpurun_out/r2_train_bf16.json; tail -3 gpurun_out/r2_train_bf16.err
echo "== x3 tests + bench"
timeout 600 python -m pytest tests/test_gpu_x3.py -m gpu -q 2>&1 | tail -3
timeout 600 python bench.py --precision fp16x3 --steps 100 --warmup 10 --no-cpu-baseline --no-frame > gpurun_out/r2_bench_fp16x3_ldg.json 2> gpurun_out/r2_bench_fp16x3_ldg.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_fp16x3_ldg.json').read().strip().splitlines()[-1])
print('fp16x3', d['value'], d['ms_per_step'], d['e2e']['value'], d.get('kernel_ms_per_step'), d['roofline']['frac'], d.get('parity'))
PY
