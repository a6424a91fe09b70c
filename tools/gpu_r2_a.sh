#!/bin/bash
# round-2 first GPU pass: new tests verbosely, then the whole gpu suite, smoke, and the two bench lines
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv | tail -1
echo "== x3 + graph + new parity tests"
timeout 600 python -m pytest tests/test_gpu_x3.py tests/test_gpu_graph.py -m gpu -q -x -s 2>&1 | grep -v "^$" | tail -60 > gpurun_out/r2_x3_tests.log; tail -45 gpurun_out/r2_x3_tests.log
echo "== full gpu suite"
timeout 1500 python -m pytest tests -m gpu -q -rP 2>&1 > gpurun_out/r2_gpu_suite.log; tail -8 gpurun_out/r2_gpu_suite.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -8
echo "== bench bf16"; timeout 600 python bench.py > gpurun_out/r2_bench_bf16.json 2> gpurun_out/r2_bench_bf16.err; tail -2 gpurun_out/r2_bench_bf16.err
echo "== bench fp16x3"; timeout 600 python bench.py --precision fp16x3 --no-cpu-baseline --no-frame > gpurun_out/r2_bench_fp16x3.json 2> gpurun_out/r2_bench_fp16x3.err; tail -2 gpurun_out/r2_bench_fp16x3.err
python - <<'PY'
import json
for n in ("bf16","fp16x3"):
    try:
        d=json.load(open(f"gpurun_out/r2_bench_{n}.json"))
        print(n,{k:d[k] for k in ("value","ms_per_step","e2e","kernel_ms","gpu_launches","clocks","parity")})
        print(n,d["roofline"])
        print(n,"frame",d.get("frame"),"cpu",d.get("cpu_baseline"))
    except Exception as e: print(n,"failed",e)
PY
