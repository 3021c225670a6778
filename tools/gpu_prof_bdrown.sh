timeout 300 python tools/bench_kernels.py --what bdrownC --reps 3 > gpurun_out/bdC_plain.log 2>&1 || { tail -3 gpurun_out/bdC_plain.log; exit 1; }
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_bdrown_smem' -s 1 -c 1 -o gpurun_out/prof_bdrownC_r02 -f python tools/bench_kernels.py --what bdrownC --reps 3 > gpurun_out/bdC_ncu.log 2>&1
ls -la gpurun_out/prof_bdrownC_r02.ncu-rep
