#!/bin/bash
# usage: scripts/gpurun_retry.sh <logfile> <gpurun args...>   -- retries while the pod answers "busy" (exit code 3)
log=$1; shift
for try in $(seq 1 20); do
  /usr/local/graft/bin/gpurun "$@" > "$log" 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 90
done
exit 3
