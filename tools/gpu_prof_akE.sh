python tools/bench_akima_E.py > gpurun_out/plain_akE.log 2>&1 || { tail -5 gpurun_out/plain_akE.log; exit 1; }
tail -2 gpurun_out/plain_akE.log
ncu --set full --clock-control none --import-source on -k regex:'k_slopes_tiled|k_akima_eval|k_akima_needed' -s 3 -c 3 -o gpurun_out/prof_akE -f python tools/bench_akima_E.py > gpurun_out/ncu_akE.log 2>&1
tail -2 gpurun_out/ncu_akE.log
