#!/bin/bash
# tools/gpurun_retry_n.sh <gpus> <logfile> <timeout> <command...>: like gpurun_retry.sh with --gpus
g=$1; shift; log=$1; shift; to=$1; shift
for i in $(seq 1 40); do
  gpurun --gpus $g --timeout $to -- "$@" > $log 2>&1
  rc=$?
  if [ $rc -ne 3 ] && ! grep -q "status=transient" $log; then exit $rc; fi
  sleep 45
done
