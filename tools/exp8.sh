tag=$1; shift
out=gpurun_out/${tag}_exp.txt
for v in "$@"; do
  echo "=== $v" >> $out
  CHUB_NOCHECK=1 CHUB_LIB=cholupdates_b200/lib/variants/$v.so timeout 120 python tools/time_update.py 16384 30 >> $out 2>&1
done
