tag=$1
export CHUB_NOCHECK=1
for v in workersonly bothfree chainfree; do
  echo "=== $v" >> gpurun_out/${tag}_exp.txt
  CHUB_LIB=cholupdates_b200/lib/variants/$v.so timeout 120 python tools/time_update.py 16384 30 >> gpurun_out/${tag}_exp.txt 2>&1
done
CHUB_PROFILE=1 timeout 120 python tools/df_profile.py 16384 C > gpurun_out/${tag}_phase.txt 2>&1
CHUB_LIB=cholupdates_b200/lib/variants/workersonly.so CHUB_PROFILE=1 timeout 120 python tools/df_profile.py 16384 C > gpurun_out/${tag}_phase_workersonly.txt 2>&1
