#!/bin/bash
# DRAM traffic of the dominant kernel at the bench configuration (batch 128): one launch, dram bytes only
mkdir -p gpurun_out
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:conv_c8_kernel -s 40 -c 60 --csv \
   --log-file gpurun_out/traffic_b128.csv python bench.py --steps 1 --warmup 3 --batch 128 > gpurun_out/ncu_traffic.log 2>&1
tail -1 gpurun_out/ncu_traffic.log | cut -c1-120; wc -l gpurun_out/traffic_b128.csv
