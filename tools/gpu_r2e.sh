set -x
timeout 600 python -m pytest tests/test_gpu_tma.py tests/test_gpu_fullsize.py -q -x > gpurun_out/r2e_pytest.log 2>&1; echo pytest_rc=$?; tail -4 gpurun_out/r2e_pytest.log
for spc in 0 148; do SOSIE_GRP_SPC=$spc timeout 300 python tools/bench_kernels.py --tma 0 2>&1 | tail -1; done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_apply' -c 8 --csv --log-file gpurun_out/r2e_launches.csv python tools/bench_kernels.py --tma 0 --reps 2 > gpurun_out/r2e_l.log 2>&1
grep -E "k_apply" gpurun_out/r2e_launches.csv | cut -c1-200 | tail -8
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_apply_grp' -s 1 -c 1 -o gpurun_out/prof_applyD_r02b -f python tools/bench_kernels.py --tma 0 --reps 2 > gpurun_out/r2e_ncu.log 2>&1
