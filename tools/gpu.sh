#!/bin/bash
# One parameterised GPU-side script (run under gpurun): each step in its own process with its own timeout, logs and
# ncu outputs under gpurun_out/. Steps are given as arguments and run in order:
#
#   tests[:EXPR]        pytest -m gpu (-k EXPR)                      -> gpurun_out/pytest_gpu.log
#   smoke               __graft_entry__.smoke()                      -> gpurun_out/smoke.log
#   bench[:ARGS]        python bench.py ARGS (comma = space)         -> gpurun_out/bench_<ARGS>.log (+ class table .md)
#   dp:N[:ARGS]         torchrun --nproc-per-node N bench.py --gpus N ARGS
#   launches[:ARGS]     ncu launch list (gpu__time_duration) of bench.py --steps 1 ARGS   -> gpurun_out/launches_<ARGS>.csv
#   traffic:REGEX[:ARGS]  dram bytes + duration per launch of kernels matching REGEX       -> gpurun_out/traffic_<..>.csv
#   full:REGEX[:ARGS]   ncu --set full --import-source on of kernels matching REGEX (-c 3) -> gpurun_out/full_<..>.ncu-rep
#   all[:ARGS]          light ncu capture (speed-of-light, memory, tensor pipe, DRAM bytes) of EVERY launch of one forward
#                                                                    -> gpurun_out/all_<ARGS>.ncu-rep (tools/ncu_summary.py)
#   precision[:B,H,W]  tools/precision_report.py: error of every precision mode vs the oracle -> gpurun_out/precision.md
#   probe[:CASES]       SE_TC_DEBUG role timers of single layers (tools/tc_probe.py)
#   env:K=V / unset:K   export K=V / unset K for the following steps
#
#   gpurun -- 'bash tools/gpu.sh tests bench bench:--size,512 launches:--batch,32'
mkdir -p gpurun_out
tag() { echo "$1" | tr -c 'A-Za-z0-9_.\n' '_' | sed 's/__*/_/g; s/^_//; s/_$//'; }
for step in "$@"; do
  kind=${step%%:*}; rest=""; [ "$step" != "$kind" ] && rest=${step#*:}
  echo "=================== $step"
  case $kind in
    env) export "$rest" ;;
    unset) unset "$rest" ;;
    tests)
      if [ -n "$rest" ]; then K=(-k "$rest"); else K=(); fi
      timeout 1500 python -m pytest tests -q -m gpu --timeout 900 "${K[@]}" > gpurun_out/pytest_gpu.log 2>&1; tail -${TAIL:-40} gpurun_out/pytest_gpu.log ;;
    smoke) ( timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -5 ) > gpurun_out/smoke.log 2>&1; cat gpurun_out/smoke.log ;;
    bench)
      args=$(echo "$rest" | tr ',' ' '); t=$(tag "bench_$rest")
      ( timeout 900 python bench.py $args --classes-out gpurun_out/${t}_classes.md 2>&1 | tail -2 ) > gpurun_out/$t.log 2>&1
      tail -1 gpurun_out/$t.log | python tools/bench_summary.py ;;
    dp)
      n=${rest%%:*}; a=""; [ "$rest" != "$n" ] && a=$(echo "${rest#*:}" | tr ',' ' '); t=$(tag "bench_dp${n}_$a")
      ( timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $n $a 2>&1 | tail -3 ) > gpurun_out/$t.log 2>&1
      tail -1 gpurun_out/$t.log | python tools/bench_summary.py ;;
    launches)
      args=$(echo "$rest" | tr ',' ' '); t=$(tag "launches_$rest")
      timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/$t.csv \
        python bench.py --steps 1 --warmup 3 --no-latency $args > gpurun_out/$t.log 2>&1
      wc -l gpurun_out/$t.csv ;;
    traffic)
      rx=${rest%%:*}; a=""; [ "$rest" != "$rx" ] && a=$(echo "${rest#*:}" | tr ',' ' '); t=$(tag "traffic_${rx}_$a")
      timeout 1200 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:$rx -s ${SKIP:-110} -c ${COUNT:-110} --csv \
        --log-file gpurun_out/$t.csv python bench.py --steps 1 --warmup 3 --no-latency $a > gpurun_out/$t.log 2>&1
      wc -l gpurun_out/$t.csv ;;
    full)
      rx=${rest%%:*}; a=""; [ "$rest" != "$rx" ] && a=$(echo "${rest#*:}" | tr ',' ' '); t=$(tag "full_${rx}_$a")
      timeout 1500 ncu --set full --clock-control none --import-source on --kernel-name-base ${NCU_BASE:-function} -k "regex:$rx" -s ${SKIP:-60} -c ${COUNT:-3} -o gpurun_out/$t -f \
        python bench.py --steps 1 --warmup 3 --no-latency $a > gpurun_out/$t.log 2>&1
      ls -la gpurun_out/$t.ncu-rep ;;
    all)
      args=$(echo "$rest" | tr ',' ' '); t=$(tag "all_$rest")
      timeout 1500 ncu --section SpeedOfLight --section LaunchStats --section MemoryWorkloadAnalysis --section Occupancy \
        --metrics dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_tensor.sum \
        --clock-control none -s ${SKIP:-170} -c ${COUNT:-90} -o gpurun_out/$t -f python bench.py --steps 1 --warmup 3 --no-latency $args > gpurun_out/$t.log 2>&1
      ls -la gpurun_out/$t.ncu-rep ;;
    precision)
      a=$(echo "$rest" | tr ',' ' ')
      timeout 900 python tools/precision_report.py $a > gpurun_out/precision.md 2> gpurun_out/precision.err; cat gpurun_out/precision.md; tail -3 gpurun_out/precision.err ;;
    probe)
      ( SE_TC_DEBUG=1 SE_PROBE_CASES=$rest PB=${PB:-32} timeout 300 python tools/tc_probe.py 2>&1 | grep -E "^==|^\[tc\]|^\[c8\]" ) > gpurun_out/probe.log 2>&1
      awk '/^==/{n=$0; c=0} /^\[(tc|c8)\]/{c++; if(c==1) print n "  " $0}' gpurun_out/probe.log | cut -c1-360 | awk '!seen[$0]++' ;;
    *) echo "unknown step $step" ;;
  esac
done
