#!/bin/bash
# Standard GPU round: parity tests, smoke, bench, ncu launch list. Logs -> gpurun_out/.
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -q -m gpu --timeout 600 -x 2>&1 | tail -25 ) > gpurun_out/pytest_gpu.log 2>&1
( timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -5 ) > gpurun_out/smoke.log 2>&1
( timeout 600 python bench.py --steps 10 --warmup 3 2>&1 | tail -3 ) > gpurun_out/bench.log 2>&1
( timeout 300 python bench.py --dtype fp32 --steps 3 --warmup 3 2>&1 | tail -3 ) > gpurun_out/bench_fp32.log 2>&1
if [ "$1" == "ncu" ]; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches.csv \
     python bench.py --steps 1 --warmup 3 --batch 32 > gpurun_out/ncu_bench.log 2>&1
fi
echo "== pytest"; cat gpurun_out/pytest_gpu.log; echo "== smoke"; cat gpurun_out/smoke.log; echo "== bench"; cat gpurun_out/bench.log; echo "== bench fp32"; cat gpurun_out/bench_fp32.log
