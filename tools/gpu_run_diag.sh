#!/bin/bash
# First-contact GPU script: each group in its own process with its own timeout; logs under gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/diag.log 2>&1
for g in "$@"; do
  timeout 300 python tools/gpu_diag.py $g >> gpurun_out/diag.log 2>&1
  echo "--- group $g exit $?" >> gpurun_out/diag.log
done
tail -c 6000 gpurun_out/diag.log
