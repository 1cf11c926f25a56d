#!/bin/bash
# A/B of library builds on one box: tools/ab_update.sh <tag> [lib ...]   (CHUB_LIB picks the build)
tag=$1; shift
for lib in "$@"; do
  name=$(basename $lib .so)
  echo "=== $name" >> gpurun_out/${tag}_ab.txt
  CHUB_LIB=$lib timeout 120 python tools/time_update.py 16384 30 >> gpurun_out/${tag}_ab.txt 2>&1
  echo "rc=$?" >> gpurun_out/${tag}_ab.txt
done
