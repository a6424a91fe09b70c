#!/bin/bash
# multi-GPU pass: NCCL tests (sharded == single, graph with all_gather) and the bench line (weak value + strong key)
N=${1:-2}
mkdir -p gpurun_out
if [ "$N" = "2-tests" ]; then
  echo "== multi-GPU tests"; timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q 2>&1 | tail -3
  exit 0
fi
echo "== bench N=$N"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 100 --warmup 10 > gpurun_out/r2_scale_n$N.json 2> gpurun_out/r2_scale_n$N.err
python - $N <<'PY'
import json,sys
n=sys.argv[1]
try:
    d=json.loads(open(f'gpurun_out/r2_scale_n{n}.json').read().strip().splitlines()[-1])
    print(n, 'rays/s', f"{d['value']:.4g}", 'ms/step', round(d['ms_per_step'],4), 'e2e', f"{d['e2e']['value']:.4g}", 'frame_ms', round(d['frame']['ms'],2), 'frac', round(d['roofline']['frac'],4), d['clocks'])
    print('strong', d['strong_scaling'])
except Exception as e: print(n,'failed',e); print(open(f'gpurun_out/r2_scale_n{n}.err').read()[-2500:])
PY
echo "== data-parallel training step N=$N"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 tools/train_bench_multi.py > gpurun_out/r2_train_multi_n$N.json 2> gpurun_out/r2_train_multi_n$N.err; tail -1 gpurun_out/r2_train_multi_n$N.json; tail -2 gpurun_out/r2_train_multi_n$N.err | cut -c1-300
