set -x
python tools/bench_drown_big.py 2560 1280 16 100 50 > gpurun_out/plain_F.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'k_sml_pass_int' -s 20 -c 1 -o gpurun_out/prof_sml_r01f -f python tools/bench_drown_big.py 2560 1280 16 100 50 > gpurun_out/ncu_F.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_bdl_level|k_bdl_sweep' -s 230 -c 2 -o gpurun_out/prof_lvl_r01f -f python tools/bench_drown_big.py 2560 1280 16 100 50 > gpurun_out/ncu_F2.log 2>&1
python tools/bench_akima_E.py > gpurun_out/plain_AK.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'k_slopes_tiled|k_akima_eval' -s 2 -c 2 -o gpurun_out/prof_akima_r01f -f python tools/bench_akima_E.py > gpurun_out/ncu_AK.log 2>&1
python tools/bench_mapC.py C > gpurun_out/plain_mapC.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'k_twisted_band' -s 2 -c 1 -o gpurun_out/prof_twisted_r01f -f python tools/bench_mapC.py C > gpurun_out/ncu_tw.log 2>&1
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo rc=$?
ls gpurun_out/*r01f*
