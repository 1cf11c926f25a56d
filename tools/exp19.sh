tag=$1
out=gpurun_out/${tag}_time_ops.txt
timeout 300 python tools/time_ops.py 4096,8192,16384,20000,32768 > $out 2>&1
echo "=== CHUB_DD_SWEEP_CM_V2=1" >> $out
CHUB_DD_SWEEP_CM_V2=1 timeout 300 python tools/time_ops.py 16384 >> $out 2>&1
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -x -q -m gpu -k "downdate or switches or blocked or 32768 or sizes" > gpurun_out/${tag}_tests.log 2>&1; tail -3 gpurun_out/${tag}_tests.log
