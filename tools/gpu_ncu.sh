#!/bin/bash
# ncu evidence: (1) launch list of one bench step, (2) full-set capture of the dominant kernel on the 96->192 layer
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches.csv \
   python bench.py --steps 1 --warmup 3 --batch 32 > gpurun_out/ncu_bench.log 2>&1
PB=32 timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_tc -c 6 -o gpurun_out/prof_conv_tc -f \
   python tools/tc_probe.py > gpurun_out/ncu_probe.log 2>&1
ls -la gpurun_out/ | tail -8
