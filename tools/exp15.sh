tag=$1
out=gpurun_out/${tag}_time_ops.txt
echo "=== pair (default)" > $out
timeout 300 python tools/time_ops.py 4096,16384,18000,20000 >> $out 2>&1
echo "=== CHUB_CM_PAIR=2" >> $out
CHUB_CM_PAIR=2 timeout 300 python tools/time_ops.py 4096,16384 >> $out 2>&1
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -x -q -m gpu > gpurun_out/${tag}_tests.log 2>&1; tail -3 gpurun_out/${tag}_tests.log
CHUB_PROFILE=2 timeout 120 python tools/df_profile.py 16384 F > gpurun_out/${tag}_phase_F.txt 2>&1
