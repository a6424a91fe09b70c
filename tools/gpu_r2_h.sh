#!/bin/bash
mkdir -p gpurun_out
echo "== full gpu suite"
timeout 1800 python -m pytest tests -m gpu -q -rP 2>&1 > gpurun_out/r2_gpu_suite.log; tail -15 gpurun_out/r2_gpu_suite.log | cut -c1-250
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -8
echo "== bench bf16"; timeout 900 python bench.py > gpurun_out/r2_bench_bf16.json 2> gpurun_out/r2_bench_bf16.err; tail -2 gpurun_out/r2_bench_bf16.err | cut -c1-300
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2_bench_bf16.json"))
print({k:d[k] for k in ("value","ms_per_step","e2e","kernel_ms","gpu_launches","clocks","parity","cpu_baseline")})
print(d["roofline"]); print(d["frame"]); print(d.get("train_step"))
PY
echo "== bench fp16x3"; timeout 900 python bench.py --precision fp16x3 --no-frame > gpurun_out/r2_bench_fp16x3.json 2> gpurun_out/r2_bench_fp16x3.err; python - <<'PY'
import json
d=json.load(open("gpurun_out/r2_bench_fp16x3.json"))
print({k:d[k] for k in ("value","ms_per_step","e2e","parity")}); print(d["roofline"]["frac"])
PY
echo "== reference arm"; timeout 900 python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | tail -1 | cut -c1-600
