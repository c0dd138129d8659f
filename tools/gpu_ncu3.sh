#!/bin/bash
# full-set ncu capture (with source correlation) of the small-N C8 kernels: stem (conv1) and conv16
mkdir -p gpurun_out
SE_PROBE_CASES=${CASES:-conv1,conv16} PB=32 timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_c8 -c ${NCAP:-4} -o gpurun_out/prof_c8_small -f \
   python tools/tc_probe.py > gpurun_out/ncu_probe3.log 2>&1
tail -3 gpurun_out/ncu_probe3.log; ls -la gpurun_out/*.ncu-rep
