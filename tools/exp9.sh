tag=$1; shift
out=gpurun_out/${tag}_exp.txt
echo "=== main" >> $out
timeout 120 python tools/time_update.py 16384 30 >> $out 2>&1
for v in "$@"; do
  echo "=== $v" >> $out
  CHUB_LIB=cholupdates_b200/lib/variants/$v.so timeout 120 python tools/time_update.py 16384 30 >> $out 2>&1
  CHUB_LIB=cholupdates_b200/lib/variants/$v.so CHUB_PROFILE=2 timeout 120 python tools/df_profile.py 16384 C > gpurun_out/${tag}_phase_$v.txt 2>&1
done
