tag=$1
timeout 900 python -m pytest tests/test_gpu_fullsize.py -x -q -m gpu -k "blocked or 32768 or above" > gpurun_out/${tag}_tests.log 2>&1; tail -15 gpurun_out/${tag}_tests.log
timeout 300 python tools/time_ops.py 30000,32768 > gpurun_out/${tag}_time_ops.txt 2>&1
