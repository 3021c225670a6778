python tools/bench_akima.py > gpurun_out/akima.json 2> gpurun_out/akima.err; echo rc=$?; tail -3 gpurun_out/akima.err; cat gpurun_out/akima.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_akima.csv python tools/bench_akima.py 16 > gpurun_out/ncuAK.log 2>&1
