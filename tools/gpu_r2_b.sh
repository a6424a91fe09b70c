#!/bin/bash
mkdir -p gpurun_out
echo "== x3 tests"
timeout 600 python -m pytest tests/test_gpu_x3.py tests/test_gpu_graph.py tests/test_gpu_training.py -m gpu -q -s 2>&1 | grep -v "^$" > gpurun_out/r2_x3_tests.log; grep -E "fp16x3|bf16x3|passed|failed|FAILED|Error" gpurun_out/r2_x3_tests.log | tail -40
echo "== host profile (bf16)"
timeout 300 python tools/host_profile.py 2>&1 | head -30
echo "== bench fp16x3 / bf16x3 (short)"
for p in fp16x3 bf16x3; do
timeout 600 python bench.py --precision $p --steps 40 --warmup 5 --no-cpu-baseline --no-frame > gpurun_out/r2_bench_$p.json 2> gpurun_out/r2_bench_$p.err; tail -2 gpurun_out/r2_bench_$p.err
python -c "
import json; d=json.load(open('gpurun_out/r2_bench_$p.json')); print('$p', d['value'], d['ms_per_step'], d['roofline']['launch_ms'], d['roofline']['frac'], d['parity'])"
done
echo "== trace fp16x3"
export MIPNERF_B200_LIB=$PWD/mipnerf_pl_b200/libmipnerf_b200.trace.so
TRACE_PRECISION=fp16x3 timeout 300 python tools/tc_trace.py pair 1 > gpurun_out/r2_trace_fp16x3_l0.txt 2>&1
grep -E "^g |^g[0-9]|slot|features|tile period|IPE|events" gpurun_out/r2_trace_fp16x3_l0.txt | head -40
TRACE_PRECISION=fp16x3 timeout 300 python tools/tc_trace.py pair 2 > gpurun_out/r2_trace_fp16x3_l1.txt 2>&1
grep -E "^g |^g[0-9]|slot|features|tile period|IPE|events" gpurun_out/r2_trace_fp16x3_l1.txt | head -20
