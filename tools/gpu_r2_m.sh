#!/bin/bash
# training step: reduce unroll + ncu of the three big kernels
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_training.py -m gpu -q 2>&1 | tail -3
timeout 600 python tools/train_bench.py --precision bf16 > gpurun_out/r2_train_bf16_fused3.json 2> gpurun_out/r2_train_bf16_fused3.err; tail -c 1300 gpurun_out/r2_train_bf16_fused3.json; tail -3 gpurun_out/r2_train_bf16_fused3.err
echo "== ncu (level kernel train mode, linear_t16, wgrad_mn)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"mlp_level_kernel|linear_t16_kernel|wgrad_mn_kernel" -s 40 -c 12 -f -o gpurun_out/r2_prof_train python tools/ncu_train_target.py bf16 4096 2>&1 | tail -2
ls -la gpurun_out/r2_prof_train.ncu-rep
