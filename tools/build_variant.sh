#!/bin/bash
# tools/build_variant.sh <name> [-DFLAG ...]: an A/B build of the library under lib/variants/ (CHUB_LIB picks it)
name=$1; shift
d=cholupdates_b200
mkdir -p $d/lib/variants
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared "$@" \
  -o $d/lib/variants/$name.so $d/csrc/api.cu $d/csrc/host.cu 2>&1 | grep -v "warning\|nsteps\|\^\|^$\|Remark\|detected during" 
echo "built $name"
