#!/bin/bash
mkdir -p gpurun_out
echo "== training tests"; timeout 900 python -m pytest tests/test_gpu_training.py -m gpu -q -s 2>&1 | grep -v "^$" | tail -40
echo "== rest of gpu suite"; timeout 1500 python -m pytest tests -m gpu -q --deselect tests/test_gpu_training.py 2>&1 | tail -6
echo "== mlp stage test output"; timeout 300 python -m pytest tests/test_gpu_tensorcore.py -m gpu -q -s -k "mlp_stage_entry and pair" 2>&1 | grep -E "raw_|passed|failed" | head
echo "== train bench"; timeout 600 python tools/train_bench.py > gpurun_out/train_bench.json 2> gpurun_out/train_bench.err; tail -3 gpurun_out/train_bench.err; cat gpurun_out/train_bench.json
echo "== bench"; timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_b15.json 2> gpurun_out/bench_b15.err; python -c "
import json; d=json.load(open('gpurun_out/bench_b15.json')); print({k:d[k] for k in ('value','ms_per_step','e2e','kernel_ms','kernel_launches','frame','clocks')}); print(d['roofline'])"; tail -3 gpurun_out/bench_b15.err
