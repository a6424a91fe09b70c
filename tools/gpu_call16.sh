#!/bin/bash
mkdir -p gpurun_out
echo "== training tests"; timeout 900 python -m pytest tests/test_gpu_training.py -m gpu -q -s 2>&1 | grep -E "case|xavier|trained_like|training losses|passed|failed|Error" | cut -c1-900 | tail -30
echo "== full gpu suite"; timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -4
