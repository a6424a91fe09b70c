#!/bin/bash
mkdir -p gpurun_out
export MIPNERF_B200_LIB=$PWD/mipnerf_pl_b200/libmipnerf_b200.trace.so
for v in shared pair; do timeout 300 python tools/tc_trace.py $v > gpurun_out/trace_$v.txt 2>&1; echo "== $v"; grep -E "events|tile period" gpurun_out/trace_$v.txt; done
unset MIPNERF_B200_LIB
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,power.limit,clocks_event_reasons.active,clocks_event_reasons.sw_power_cap,clocks_event_reasons.hw_slowdown --format=csv -lms 50 > gpurun_out/smi_during.csv &
SMI=$!
sleep 1
for v in shared pair; do MIPNERF_B200_TC_VARIANT=$v python bench.py --steps 2000 --warmup 50 --precision bf16 --no-cpu-baseline > gpurun_out/bench_long_$v.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/bench_long_$v.json')); print('$v', d['value'], d['roofline']['launch_ms'], d['roofline']['frac'], d['clocks'])"; done
kill $SMI
sort gpurun_out/smi_during.csv | uniq -c | sort -rn | head -12
