tag=$1
out=gpurun_out/${tag}_exp.txt
echo "=== main (v6, d=5)" > $out
timeout 120 python tools/time_update.py 16384 30 >> $out 2>&1
for v in v7d5 v7d6s3 v7d7s3 v6d7s3; do
  echo "=== $v" >> $out
  CHUB_LIB=cholupdates_b200/lib/variants/$v.so timeout 120 python tools/time_update.py 16384 30 >> $out 2>&1
done
