mkdir -p gpurun_out/ncu_r2
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/ncu_r2/launches_fastdb.csv python tools/fastdb_perf.py 100 20000 > gpurun_out/run4.log 2>&1
RSK_TIMING=1 python tools/fastdb_perf.py 100 20000 >> gpurun_out/run4.log 2>&1
