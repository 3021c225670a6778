set -x
for spc in 0 96 128; do SOSIE_GRP_SPC=$spc timeout 300 python tools/bench_kernels.py --tma 0 2>&1 | tail -1; done
for spc in 0 96 128 244; do SOSIE_GRP_SPC=$spc timeout 300 python tools/bench_kernels.py --tma 0 --slabs 1460 2>&1 | tail -1; done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_apply_grp' -s 1 -c 1 -o gpurun_out/prof_applyD_r02c -f python tools/bench_kernels.py --tma 0 --reps 2 > gpurun_out/r2h_ncu.log 2>&1
