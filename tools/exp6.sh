tag=$1
out=gpurun_out/${tag}_exp.txt
echo "=== main (v6 chain)" >> $out
timeout 120 python tools/time_update.py 16384 30 >> $out 2>&1
echo "=== barrier chain" >> $out
CHUB_LIB=cholupdates_b200/lib/variants/barrier.so timeout 120 python tools/time_update.py 16384 30 >> $out 2>&1
for n in 96 130 1000 4000 8200; do
  echo "=== n=$n" >> $out
  timeout 120 python tools/time_update.py $n 10 >> $out 2>&1
done
CHUB_PROFILE=2 timeout 120 python tools/df_profile.py 16384 C > gpurun_out/${tag}_phase.txt 2>&1
timeout 300 python tools/time_ops.py > gpurun_out/${tag}_time_ops.txt 2>&1
