tag=$1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${tag}_launches_20000.csv python tools/one_each.py 20000 C update > gpurun_out/${tag}_ncu.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${tag}_launches_32768.csv python tools/one_each.py 32768 C update >> gpurun_out/${tag}_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:df_kernel_rm -c 1 -o gpurun_out/${tag}_df_kernel_rm python tools/one_each.py 16384 C update >> gpurun_out/${tag}_ncu.log 2>&1
tail -3 gpurun_out/${tag}_ncu.log
