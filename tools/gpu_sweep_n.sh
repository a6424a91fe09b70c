#!/bin/bash
# BASELINE configs[4] at N GPUs (N = $1): the ray-batch sweep under torchrun, rays sharded per rank.
N=${1:-4}
mkdir -p gpurun_out
timeout 280 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29546 tools/sweep.py --steps 15 --warmup 3 > gpurun_out/sweep_n$N.json 2> gpurun_out/sweep_n$N.err
tail -2 gpurun_out/sweep_n$N.err | cut -c1-300
python -c "
import json; d=json.loads(open('gpurun_out/sweep_n$N.json').read().strip().splitlines()[-1])
for r in d['rows']: print(r['rays_per_gpu'], round(r['ms_per_step'],3), f\"{r['rays_per_s']:.4g}\", round(r['frac_of_mlp_roofline_per_gpu'],3))"
