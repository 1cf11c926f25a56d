export CHUB_NOCHECK=1
tag=$1; shift
for v in "$@"; do
  echo "=== $v" >> gpurun_out/${tag}_exp.txt
  CHUB_LIB=cholupdates_b200/lib/variants/$v.so timeout 120 python tools/time_update.py 16384 30 >> gpurun_out/${tag}_exp.txt 2>&1
done
