#!/bin/bash
# how does the streamed-weight (pair) kernel react to the depth of its weight ring?
mkdir -p gpurun_out
for st in 3 5 7; do
  echo "== stages cap $st"
  SE_C8_STAGES=$st SE_TC_DEBUG=1 SE_PROBE_CASES=conv5,conv11 PB=32 timeout 200 python tools/tc_probe.py 2>&1 | grep -E "^\[c8\]" | awk 'NR%2==0' | cut -c1-330
done
for pb in 8 16 64; do
  echo "== batch $pb"
  SE_TC_DEBUG=1 SE_PROBE_CASES=conv5 PB=$pb timeout 200 python tools/tc_probe.py 2>&1 | grep -E "^\[c8\]" | awk 'NR%2==0' | cut -c1-330
done
