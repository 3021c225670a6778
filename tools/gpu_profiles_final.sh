set -x
B="python bench.py --steps 2 --warmup 3 --no-cpu --no-others --no-e2e"
$B > gpurun_out/plain_E.log 2>&1 || { tail -5 gpurun_out/plain_E.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_E_r01c.csv $B > gpurun_out/ncu_l.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_map_easy|k_apply1|k_pack' -s 9 -c 3 -o gpurun_out/prof_E_r01c -f $B > gpurun_out/ncu_f.log 2>&1
python tools/bench_kernels.py > gpurun_out/plain_D.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'k_apply_staged' -s 1 -c 1 -o gpurun_out/prof_applyD_r01c -f python tools/bench_kernels.py > gpurun_out/ncu_d.log 2>&1
python tools/bench_drown_big.py 2560 1280 16 100 50 > gpurun_out/plain_F.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'k_sml_pass|k_bdl_level|k_bdl_sweep' -s 330 -c 3 -o gpurun_out/prof_drownF_r01c -f python tools/bench_drown_big.py 2560 1280 16 100 50 > gpurun_out/ncu_F.log 2>&1
python tools/bench_bdrown_C.py > gpurun_out/plain_C.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'k_bdrown_smem' -s 3 -c 1 -o gpurun_out/prof_bdrownC_r01c -f python tools/bench_bdrown_C.py > gpurun_out/ncu_C.log 2>&1
python tools/bench_akima.py 16 > gpurun_out/plain_AK.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'k_slopes_tiled|k_akima_eval' -s 8 -c 2 -o gpurun_out/prof_akima_r01c -f python tools/bench_akima.py 16 > gpurun_out/ncu_AK.log 2>&1
ls -la gpurun_out/*r01c*
