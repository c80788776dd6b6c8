#!/bin/bash
# usage: [GPURUN_FLAGS="--gpus 2"] scripts/gpurun_retry.sh <log file> <timeout s> <command...>
# retries while the pod answers "busy" (nothing charged) or another call of this repo is still in flight
log=$1; shift; to=$1; shift
for i in $(seq 1 200); do
  /usr/local/graft/bin/gpurun $GPURUN_FLAGS --timeout "$to" -- "$@" > "$log" 2>&1
  rc=$?
  if grep -q "status=transient" "$log" || grep -q "already running" "$log" || [ $rc -eq 3 ]; then sleep 40; continue; fi
  exit $rc
done
exit 3
