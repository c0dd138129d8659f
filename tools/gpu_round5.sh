#!/bin/bash
mkdir -p gpurun_out
( timeout 300 python tools/gpu_diag.py tc_plain tc_s2 tc_dil tc_stem tc_deconv tc_small heads 2>&1 | grep -E "^(bf16 |FAIL|===)" ) > gpurun_out/diag5.log 2>&1
( timeout 200 python tools/gpu_diag.py nets e2e 2>&1 | grep -E "^(net|e2e|FAIL)" ) >> gpurun_out/diag5.log 2>&1
( timeout 900 python -m pytest tests -q -m gpu --timeout 600 2>&1 | tail -8 ) > gpurun_out/pytest_gpu.log 2>&1
( SE_TC_DEBUG=1 timeout 300 python tools/tc_probe.py 2>&1 | grep -E "^==|^\[tc\]|^\[c8\]" ) > gpurun_out/probe.log 2>&1
( timeout 600 python bench.py --steps 10 --warmup 3 2>&1 | tail -1 ) > gpurun_out/bench.log 2>&1
echo "== diag"; cat gpurun_out/diag5.log; echo "== pytest"; cat gpurun_out/pytest_gpu.log; echo "== probe"; awk '/^==/{n=$0; c=0} /^\[(tc|c8)\]/{c++; if(c==1) print n "  " $0}' gpurun_out/probe.log | cut -c1-360 | awk '!seen[$0]++' ; echo "== bench"; cat gpurun_out/bench.log | python -c "
import sys, json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print({k: d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'], d['roofline'] and {k: d['roofline'][k] for k in ('achieved','frac','kernel_share_of_step')}, d.get('cpu_baseline',{}).get('value'))"
