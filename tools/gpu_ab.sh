#!/bin/bash
# same-box A/B of an environment switch on the role-timer probe: AB_ENV="SE_C8_NOKS1=1"
mkdir -p gpurun_out
for rep in 1 2; do
  for mode in base alt; do
    if [ $mode = alt ]; then export $AB_ENV; else unset ${AB_ENV%%=*}; fi
    echo "== $mode rep $rep"
    SE_TC_DEBUG=1 SE_PROBE_CASES=${CASES:-conv1,conv3,conv13_upsample_conv,conv15_upsample_conv,conv16,conv2_downsample} PB=32 timeout 200 python tools/tc_probe.py 2>&1 | grep -E "^\[c8\]" | awk 'NR%2==0' | sed -E 's/.*(Ci=[0-9]+ taps=[0-9]+ NT=[0-9]+).*(\| mma.*)/\1 \2/' | cut -c1-200
  done
done
