#!/bin/bash
# usage: tools/gpurun_retry.sh <logfile> <gpurun args...>   retries while the pod answers "busy" (nothing charged)
log=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun "$@" > $log 2>&1
  if ! grep -q "status=transient" $log; then exit 0; fi
  sleep 90
done
