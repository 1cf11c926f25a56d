tag=$1
echo "=== update n=16384 C" > gpurun_out/${tag}_exp.txt
timeout 120 python tools/time_update.py 16384 30 >> gpurun_out/${tag}_exp.txt 2>&1
timeout 300 python tools/time_ops.py 4096,16384,20000 > gpurun_out/${tag}_time_ops.txt 2>&1
CHUB_PROFILE=2 timeout 120 python tools/df_profile.py 16384 F > gpurun_out/${tag}_phase_F.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/${tag}_tests.log 2>&1
tail -3 gpurun_out/${tag}_tests.log
