#!/bin/bash
# tools/grun.sh <timeout-seconds> '<command>' : gpurun with retries while the pod is busy (exit code 3 = nothing charged)
T=$1; shift
for i in $(seq 1 30); do
  /usr/local/graft/bin/gpurun --timeout $T -- "$@" > /tmp/grun_last.log 2>&1
  rc=$?
  if grep -q "status=transient" /tmp/grun_last.log || [ $rc -eq 3 ]; then sleep 60; continue; fi
  break
done
cat /tmp/grun_last.log
exit $rc
