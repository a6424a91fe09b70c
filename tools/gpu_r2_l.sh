#!/bin/bash
# fused training step, second pass: single-launch packer (inference parity must hold), fp16 gradient scale, epilogue split
mkdir -p gpurun_out
echo "== training + rng tests"
timeout 900 python -m pytest tests/test_gpu_training.py tests/test_gpu_rng.py -m gpu -q 2>&1 | tail -4
echo "== tracks fp32, fused only"
timeout 600 python -m pytest tests/test_gpu_training.py -m gpu -q -s -k "tracks_fp32 and fused" 2>&1 | grep -E "per-tensor|passed|failed|Error|error" | cut -c1-1200 | tail -4
echo "== inference parity (packer changed)"
timeout 900 python -m pytest tests/test_gpu_tensorcore.py tests/test_gpu_x3.py tests/test_gpu_parity.py -m gpu -q 2>&1 | tail -4
echo "== train bench bf16 / fp16"
for P in bf16 fp16; do timeout 600 python tools/train_bench.py --precision $P > gpurun_out/r2_train_${P}_fused2.json 2> gpurun_out/r2_train_${P}_fused2.err; tail -c 1300 gpurun_out/r2_train_${P}_fused2.json; tail -3 gpurun_out/r2_train_${P}_fused2.err; done
