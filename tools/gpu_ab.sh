#!/bin/bash
# Same-box A/B of an environment switch of the library: usage  gpu_ab.sh VAR valA valB  (two alternating pairs)
mkdir -p gpurun_out
var=$1; a=$2; b=$3
for val in $a $b $a $b; do
  env $var=$val timeout 600 python bench.py --no-cpu-baseline --no-frame > gpurun_out/bench_ab_$val.json 2> gpurun_out/bench_ab_$val.err
  python -c "
import json; d=json.load(open('gpurun_out/bench_ab_$val.json')); print('$var=$val', round(d['value']), round(d['ms_per_step'],4), d['kernel_ms'], round(d['roofline']['frac'],4), round(d['roofline']['step_frac_of_roofline'],4), 'e2e', round(d['e2e']['value']))"
done
