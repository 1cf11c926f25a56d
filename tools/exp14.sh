tag=$1
out=gpurun_out/${tag}_exp.txt
echo "=== main" >> $out
timeout 120 python tools/time_update.py 16384 30 >> $out 2>&1
bash tools/exp8.sh $tag co co_neither
CHUB_PROFILE=2 timeout 120 python tools/df_profile.py 16384 C > gpurun_out/${tag}_phase.txt 2>&1
