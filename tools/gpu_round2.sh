#!/bin/bash
mkdir -p gpurun_out
( timeout 300 python tools/gpu_diag.py tc_stem tc_small tc_plain heads 2>&1 | grep -E "^(bf16|fp32|FAIL|===)" ) > gpurun_out/diag2.log 2>&1
( timeout 900 python -m pytest tests -q -m gpu --timeout 600 2>&1 | tail -15 ) > gpurun_out/pytest_gpu.log 2>&1
( SE_TC_DEBUG=1 timeout 300 python tools/tc_probe.py 2>&1 | grep -E "^==|^\[tc\]" ) > gpurun_out/probe.log 2>&1
( timeout 600 python bench.py --steps 10 --warmup 3 2>&1 | tail -1 ) > gpurun_out/bench.log 2>&1
echo "== diag"; cat gpurun_out/diag2.log; echo "== pytest"; cat gpurun_out/pytest_gpu.log; echo "== probe"; awk '/^==/{n=$0; c=0} /^\[tc\]/{c++; if(c==1) print n "  " $0}' gpurun_out/probe.log | cut -c1-330; echo "== bench"; cat gpurun_out/bench.log
