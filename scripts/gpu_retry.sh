#!/bin/bash
# usage: scripts/gpu_retry.sh <timeout_s> <command...>   (retries while the pod has no free slot)
T=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout $T -- "$@" > /tmp/gpurun_last.log 2>&1
  rc=$?
  if grep -q "status=transient" /tmp/gpurun_last.log; then sleep 45; continue; fi
  break
done
tail -150 /tmp/gpurun_last.log
exit $rc
