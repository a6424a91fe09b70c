#!/bin/bash
mkdir -p gpurun_out
for v in base hint base hint20us hint base; do
  if [ "$v" = "base" ]; then unset MIPNERF_B200_LIB; else export MIPNERF_B200_LIB=$PWD/mipnerf_pl_b200/libmipnerf_b200.$v.so; fi
  timeout 600 python bench.py --no-cpu-baseline --no-frame > gpurun_out/bench_lib_$v.json 2> gpurun_out/bench_lib_$v.err
  python -c "
import json; d=json.load(open('gpurun_out/bench_lib_$v.json')); print('$v', round(d['value']), round(d['ms_per_step'],4), d['kernel_ms'], round(d['roofline']['frac'],4), round(d['roofline']['step_frac_of_roofline'],4))"
done
