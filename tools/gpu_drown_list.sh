ncu --metrics gpu__time_duration.sum --clock-control none -c 1300 --csv --log-file gpurun_out/launches_drown_big.csv python tools/bench_drown_big.py 2560 1280 64 100 50 > gpurun_out/ncuDL.log 2>&1
tail -2 gpurun_out/ncuDL.log
