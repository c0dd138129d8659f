#!/bin/bash
# build-container side: keep asking for a GPU slot while the pod answers "busy" (exit code 3, nothing charged)
#   tools/gpurun_retry.sh <logfile> <gpurun args...>
log=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun "$@" > "$log" 2>&1
  rc=$?
  [ $rc -ne 3 ] && exit $rc
  sleep 90
done
exit 3
