tag=$1
out=gpurun_out/${tag}_time_ops.txt
timeout 300 python tools/time_ops.py 1000,4096,16384,20000,32768 > $out 2>&1
timeout 120 python tools/time_update.py 16384 30 >> $out 2>&1
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -x -q -m gpu > gpurun_out/${tag}_tests.log 2>&1; tail -5 gpurun_out/${tag}_tests.log
CHUB_PROFILE=2 timeout 120 python tools/df_profile.py 16384 C > gpurun_out/${tag}_phase.txt 2>&1
