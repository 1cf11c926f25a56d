tag=$1
bash tools/exp8.sh $tag co co_noband co_notma co_neither
timeout 300 python tools/time_ops.py 20000,32768 > gpurun_out/${tag}_time_ops.txt 2>&1
timeout 300 python -m pytest tests/test_gpu_fullsize.py -x -q -m gpu > gpurun_out/${tag}_tests.log 2>&1; tail -2 gpurun_out/${tag}_tests.log
