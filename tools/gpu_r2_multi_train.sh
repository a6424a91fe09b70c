#!/bin/bash
N=${1:-8}
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 tools/train_bench_multi.py > gpurun_out/r2_train_multi_n$N.json 2> gpurun_out/r2_train_multi_n$N.err; tail -1 gpurun_out/r2_train_multi_n$N.json; tail -2 gpurun_out/r2_train_multi_n$N.err | cut -c1-300
