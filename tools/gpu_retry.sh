#!/bin/bash
# retry a gpurun call while the pool answers "transient" (nothing charged): tools/gpu_retry.sh <timeout> <gpus> <command...>
T=$1; G=$2; shift 2
for i in $(seq 1 40); do
  if [ "$G" = "1" ]; then OUT=$(/usr/local/graft/bin/gpurun --timeout $T -- "$@" 2>&1); else OUT=$(/usr/local/graft/bin/gpurun --gpus $G --timeout $T -- "$@" 2>&1); fi
  if echo "$OUT" | grep -q "status=transient"; then sleep 60; continue; fi
  echo "$OUT" | tail -60; exit 0
done
echo "gave up after 40 transient answers"
