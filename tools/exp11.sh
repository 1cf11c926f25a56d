tag=$1
timeout 300 python tools/time_ops.py 1000,4096,16384,20000 > gpurun_out/${tag}_time_ops.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -x -q -m gpu -k "downdate or Downdate or dd or indefinite or ill" > gpurun_out/${tag}_tests.log 2>&1
tail -3 gpurun_out/${tag}_tests.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${tag}_launches.csv python tools/one_each.py 16384 C downdate > gpurun_out/${tag}_ncu.log 2>&1
