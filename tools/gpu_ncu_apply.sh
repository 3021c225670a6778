set -x
python tools/bench_kernels.py --slabs 146 --reps 2 > gpurun_out/plainD.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'k_apply' -s 1 -c 1 -o gpurun_out/prof_applyD_r01 python tools/bench_kernels.py --slabs 146 --reps 2 > gpurun_out/ncuD.log 2>&1
tail -3 gpurun_out/ncuD.log
