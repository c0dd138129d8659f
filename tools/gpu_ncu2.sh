#!/bin/bash
mkdir -p gpurun_out
SE_PROBE_CASES=conv16,conv1,conv5 PB=32 timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_c8 -c 6 -o gpurun_out/prof_conv_c8 -f \
   python tools/tc_probe.py > gpurun_out/ncu_probe2.log 2>&1
tail -3 gpurun_out/ncu_probe2.log; ls -la gpurun_out/*.ncu-rep
