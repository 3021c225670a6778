set -x
timeout 300 python tools/bench_kernels.py --tma 0 > gpurun_out/r2c_plain.log 2>&1 || { tail -5 gpurun_out/r2c_plain.log; exit 1; }
tail -2 gpurun_out/r2c_plain.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_apply_grp' -s 1 -c 1 -o gpurun_out/prof_applyD_r02a -f python tools/bench_kernels.py --tma 0 --reps 2 > gpurun_out/r2c_ncu.log 2>&1
tail -3 gpurun_out/r2c_ncu.log
ls -la gpurun_out/*.ncu-rep
