#!/bin/bash
mkdir -p gpurun_out
echo skip
N=${1:-2}
echo "== bench N=1"; timeout 300 python bench.py --no-cpu-baseline > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err || tail -5 gpurun_out/bench_n1.err
echo "== bench N=$N"; timeout 180 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err || tail -8 gpurun_out/bench_n$N.err
python - $N <<'PY'
import json,sys
for n in (1,int(sys.argv[1])):
    try:
        d=json.loads(open(f'gpurun_out/bench_n{n}.json').read().strip().splitlines()[-1])
        print(n, 'rays/s', f"{d['value']:.4g}", 'ms/step', round(d['ms_per_step'],4), 'e2e', f"{d['e2e']['value']:.4g}", 'frame_ms', round(d['frame']['ms'],2), 'frac', round(d['roofline']['frac'],4), d['clocks'])
    except Exception as e: print(n,'failed',e)
PY
