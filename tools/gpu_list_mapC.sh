python tools/bench_mapC.py C > gpurun_out/plainC.log 2>&1 && cat gpurun_out/plainC.log && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_mapC.csv python tools/bench_mapC.py C > gpurun_out/ncuC.log 2>&1
