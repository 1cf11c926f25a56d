#!/bin/bash
# tools/gpurun_retry.sh <logfile> <timeout> <command...>: retry while the pool answers "busy" (exit 3)
log=$1; shift; to=$1; shift
for i in $(seq 1 40); do
  gpurun --timeout $to -- "$@" > $log 2>&1
  rc=$?
  if [ $rc -ne 3 ] && ! grep -q "status=transient" $log; then exit $rc; fi
  sleep 45
done
