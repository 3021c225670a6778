set -x
python tools/bench_kernels.py --what bdrownC --reps 2 > gpurun_out/plainB.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'k_bdrown_smem' -s 1 -c 1 -o gpurun_out/prof_bdrownC_r01 python tools/bench_kernels.py --what bdrownC --reps 2 > gpurun_out/ncuB.log 2>&1
tail -2 gpurun_out/ncuB.log
