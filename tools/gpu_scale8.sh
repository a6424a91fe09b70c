#!/bin/bash
# 8-GPU evidence: the bench line at N=8 and the batch sweep (BASELINE configs[4]) under torchrun.
mkdir -p gpurun_out
bash tools/gpu_scale.sh 8
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29544 tools/sweep.py --steps 20 > gpurun_out/sweep_n8.json 2> gpurun_out/sweep_n8.err; tail -2 gpurun_out/sweep_n8.err | cut -c1-300; python -c "
import json; d=json.loads(open("gpurun_out/sweep_n8.json").read().strip().splitlines()[-1])
for r in d['rows']: print(r['rays_per_gpu'], round(r['ms_per_step'],3), f\"{r['rays_per_s']:.4g}\", round(r['frac_of_mlp_roofline_per_gpu'],3))"
