#!/bin/bash
# fused training step bring-up
mkdir -p gpurun_out
echo "== wgrad unit + forward identity"
timeout 600 python -m pytest tests/test_gpu_training.py -m gpu -q -x -k "wgrad_tc_matches or fused_training_forward" 2>&1 | tail -8
echo "== tracks fp32 (prints per-tensor errors)"
timeout 600 python -m pytest tests/test_gpu_training.py -m gpu -q -s -k "tracks_fp32" 2>&1 | grep -E "per-tensor|passed|failed|Error|error" | cut -c1-900 | tail -14
echo "== all training tests + rng"
timeout 900 python -m pytest tests/test_gpu_training.py tests/test_gpu_rng.py -m gpu -q 2>&1 | tail -6
echo "== train bench bf16 fused / unfused"
timeout 600 python tools/train_bench.py --precision bf16 > gpurun_out/r2_train_bf16_fused.json 2> gpurun_out/r2_train_bf16_fused.err; tail -c 1300 gpurun_out/r2_train_bf16_fused.json; tail -3 gpurun_out/r2_train_bf16_fused.err
