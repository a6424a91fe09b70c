#!/bin/bash
mkdir -p gpurun_out
echo "== training tests"; timeout 600 python -m pytest tests/test_gpu_training.py -m gpu -q 2>&1 | tail -3
for prec in fp32 bf16; do
  timeout 300 python tools/train_bench.py --precision $prec > gpurun_out/train_bench_$prec.json 2> gpurun_out/train_bench_$prec.err
  python -c "
import json; d=json.load(open('gpurun_out/train_bench_$prec.json')); print('$prec', round(d['ms_per_step'],2), round(d['rays_per_s']), {k:v for k,v in d['kernel_ms_per_step'].items() if v>1})"
done
MIPNERF_B200_WGRAD_TC=1 timeout 300 python tools/train_bench.py --precision bf16 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('bf16+wgrad_tc', round(d['ms_per_step'],2), {k:v for k,v in d['kernel_ms_per_step'].items() if v>1})"
