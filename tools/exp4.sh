tag=$1
out=gpurun_out/${tag}_exp.txt
export CHUB_NOCHECK=1
for pad in 0 16 32 64 528; do
  echo "=== ldpad=$pad" >> $out
  CHUB_LDPAD=$pad timeout 120 python tools/time_update.py 16384 20 >> $out 2>&1
done
echo "=== workersonly ldpad=32" >> $out
CHUB_LDPAD=32 CHUB_LIB=cholupdates_b200/lib/variants/workersonly.so timeout 120 python tools/time_update.py 16384 20 >> $out 2>&1
CHUB_PROFILE=1 timeout 120 python tools/df_profile.py 16384 C > gpurun_out/${tag}_phase.txt 2>&1
CHUB_LIB=cholupdates_b200/lib/variants/workersonly.so CHUB_PROFILE=1 timeout 120 python tools/df_profile.py 16384 C > gpurun_out/${tag}_phase_workersonly.txt 2>&1
CHUB_LIB=cholupdates_b200/lib/variants/chainfree.so CHUB_PROFILE=1 timeout 120 python tools/df_profile.py 16384 C > gpurun_out/${tag}_phase_chainfree.txt 2>&1
