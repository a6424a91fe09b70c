#!/bin/bash
mkdir -p gpurun_out
echo "== tc_diag sw64"; timeout 300 python tools/tc_diag.py > gpurun_out/tc_diag_sw64.txt 2>&1; cat gpurun_out/tc_diag_sw64.txt | tail
echo "== pytest gpu"; timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -15 | tee gpurun_out/pytest_gpu2.txt
