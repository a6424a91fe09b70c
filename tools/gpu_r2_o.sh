#!/bin/bash
# training step, fourth pass: streaming narrow-head wgrad, chunked pack; level-kernel dump experiments
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_training.py -m gpu -q 2>&1 | tail -3
timeout 600 python tools/train_bench.py --precision bf16 > gpurun_out/r2_train_bf16_fused5.json 2> gpurun_out/r2_train_bf16_fused5.err; tail -c 1300 gpurun_out/r2_train_bf16_fused5.json; tail -3 gpurun_out/r2_train_bf16_fused5.err
for V in nostore novdump; do
  echo "== experiment $V (timing only)"
  MIPNERF_B200_LIB=$PWD/mipnerf_pl_b200/libmipnerf_b200.$V.so timeout 600 python tools/train_bench.py --precision bf16 2>&1 | python -c "
import sys, json
for line in sys.stdin:
    if line.startswith('{'):
        d=json.loads(line); print(d['ms_per_step'], d['kernel_ms_per_step'].get('mlp_level_tc'))
"
done
