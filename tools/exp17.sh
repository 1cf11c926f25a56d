tag=$1
python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; tail -c 600 gpurun_out/${tag}_bench.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${tag}_launches_F.csv python tools/one_each.py 16384 F update,downdate > gpurun_out/${tag}_ncu.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${tag}_launches_32768_F.csv python tools/one_each.py 32768 F update,downdate >> gpurun_out/${tag}_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:df_kernel_cm2 -c 1 -o gpurun_out/${tag}_df_kernel_cm2 python tools/one_each.py 16384 F update >> gpurun_out/${tag}_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rect_apply_cm -c 1 -o gpurun_out/${tag}_rect_apply_cm python tools/one_each.py 32768 F update >> gpurun_out/${tag}_ncu.log 2>&1
