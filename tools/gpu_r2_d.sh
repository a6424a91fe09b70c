#!/bin/bash
mkdir -p gpurun_out
echo "== x3 with rolled epilogue"
for p in fp16x3; do
timeout 600 python bench.py --precision $p --steps 40 --warmup 5 --no-cpu-baseline --no-frame > gpurun_out/r2_bench_$p.json 2> gpurun_out/r2_bench_$p.err; tail -2 gpurun_out/r2_bench_$p.err
python -c "
import json; d=json.load(open('gpurun_out/r2_bench_$p.json')); print('$p', d['value'], d['ms_per_step'], d['roofline']['launch_ms'], d['roofline']['frac'], d['parity']['trained_like'])"
done
echo "== bf16 A/B: unrolled vs rolled epilogue"
for v in "" rolled; do
  if [ -n "$v" ]; then export MIPNERF_B200_LIB=$PWD/mipnerf_pl_b200/libmipnerf_b200.$v.so; else unset MIPNERF_B200_LIB; fi
  timeout 600 python bench.py --steps 100 --warmup 10 --no-cpu-baseline --no-frame --no-parity > gpurun_out/r2_bench_ab_$v.json 2> gpurun_out/r2_bench_ab_$v.err; tail -2 gpurun_out/r2_bench_ab_$v.err
  python -c "
import json; d=json.load(open('gpurun_out/r2_bench_ab_$v.json')); print('variant [$v]', d['value'], d['ms_per_step'], d['roofline']['launch_ms'], d['roofline']['frac'])"
done
unset MIPNERF_B200_LIB
export MIPNERF_B200_TC_VARIANT=v3
echo "== v3 selftests (mlp stage entry first)"
timeout 300 python -m pytest tests/test_gpu_tensorcore.py -m gpu -q -x -s -k "v3 and (mlp_stage_entry or emulated)" 2>&1 | grep -v "^$" | tail -25
echo "== v3 all tensorcore tests"
timeout 600 python -m pytest tests/test_gpu_tensorcore.py -m gpu -q -x -k "v3" 2>&1 | tail -8
echo "== bench v3"
timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-frame > gpurun_out/r2_bench_v3.json 2> gpurun_out/r2_bench_v3.err; tail -3 gpurun_out/r2_bench_v3.err
python -c "
import json; d=json.load(open('gpurun_out/r2_bench_v3.json')); print('v3', d['value'], d['ms_per_step'], d['e2e']['value'], d['kernel_ms'], d['roofline']['launch_ms'], d['roofline']['frac'], d['parity'])"
