#!/bin/bash
mkdir -p gpurun_out
echo "== gpu suite"; timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6
echo "== mlp stage test output"; timeout 300 python -m pytest tests/test_gpu_tensorcore.py -m gpu -q -s -k "mlp_stage and pair" 2>&1 | grep -E "raw_|passed|failed" | head
echo "== ray kernel bw"; timeout 300 python tools/ray_kernel_bw.py > gpurun_out/ray_kernel_bw.json 2> gpurun_out/ray_kernel_bw.err; tail -3 gpurun_out/ray_kernel_bw.err; python -c "
import json; d=json.load(open('gpurun_out/ray_kernel_bw.json'))
for r in d['rows']: print(r['stage'], r['kernel_ms_per_call'], r['achieved_GBps'], r['frac_of_hbm_peak'])"
echo "== sweep"; timeout 300 python tools/sweep.py > gpurun_out/sweep_n1.json 2> gpurun_out/sweep_n1.err; tail -2 gpurun_out/sweep_n1.err; cat gpurun_out/sweep_n1.json
timeout 300 python tools/sweep.py --multiscale > gpurun_out/sweep_n1_multiscale.json 2> gpurun_out/sweep_n1_ms.err; cat gpurun_out/sweep_n1_multiscale.json
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "== bench"; timeout 600 python bench.py > gpurun_out/bench_b14.json 2> gpurun_out/bench_b14.err; python -c "
import json; d=json.load(open('gpurun_out/bench_b14.json')); print({k:d[k] for k in ('value','ms_per_step','e2e','kernel_ms','kernel_launches','frame','cpu_baseline','clocks')}); print(d['roofline'])"; tail -3 gpurun_out/bench_b14.err
