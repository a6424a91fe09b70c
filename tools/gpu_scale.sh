#!/bin/bash
mkdir -p gpurun_out
for N in "$@"; do
  if [ "$N" = "1" ]; then timeout 200 python bench.py --no-cpu-baseline > gpurun_out/scale_n$N.json 2> gpurun_out/scale_n$N.err
  else timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N > gpurun_out/scale_n$N.json 2> gpurun_out/scale_n$N.err; fi
  python - $N <<'PY'
import json,sys
n=sys.argv[1]
try:
    d=json.loads(open(f'gpurun_out/scale_n{n}.json').read().strip().splitlines()[-1])
    print(n, 'rays/s', f"{d['value']:.4g}", 'ms/step', round(d['ms_per_step'],4), 'e2e', f"{d['e2e']['value']:.4g}", 'frame_ms', round(d['frame']['ms'],2), 'frac', round(d['roofline']['frac'],4), d['clocks'])
except Exception as e: print(n,'failed',e); print(open(f'gpurun_out/scale_n{n}.err').read()[-1500:])
PY
done
