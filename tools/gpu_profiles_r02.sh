set -x
# 1. E step: launch list + full ncu of k_map_easy / k_apply1 (json for roofline.traffic)
B="python bench.py --steps 2 --warmup 3 --no-cpu --no-others --no-e2e"
$B > gpurun_out/f1_plain_E.log 2>&1 || { tail -5 gpurun_out/f1_plain_E.log; exit 1; }
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_E_r02.csv $B > gpurun_out/f1_l.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_map_easy|k_apply1' -s 6 -c 2 -o gpurun_out/prof_E_r02 -f $B > gpurun_out/f1_f.log 2>&1
# 2. D batched apply, 584 slabs: full ncu (json for frac_hbm_dram)
timeout 300 python tools/bench_kernels.py --tma 0 --slabs 584 > gpurun_out/f1_plain_D.log 2>&1; tail -1 gpurun_out/f1_plain_D.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_apply_grp' -s 1 -c 1 -o gpurun_out/prof_applyD_r02 -f python tools/bench_kernels.py --tma 0 --slabs 584 --reps 2 > gpurun_out/f1_d.log 2>&1
ls -la gpurun_out/*.ncu-rep
