#!/bin/bash
# usage: tools/gpurun_retry.sh <timeout-seconds> <command...>: retries while the pod answers "transient" (nothing charged)
T=$1; shift
for i in $(seq 1 20); do
  /usr/local/graft/bin/gpurun --timeout $T -- "$@" > gpurun_out/.retry.log 2>&1
  if ! grep -q "status=transient" gpurun_out/.retry.log; then cat gpurun_out/.retry.log; exit 0; fi
  sleep 150
done
cat gpurun_out/.retry.log
