timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2p_l.csv python tools/bench_akima.py 4 > gpurun_out/r2p.log 2>&1
tail -2 gpurun_out/r2p.log
