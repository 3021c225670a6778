set -x
timeout 600 python -m pytest tests/test_gpu_tma.py tests/test_gpu_fullsize.py tests/test_gpu_queue.py -q -x > gpurun_out/r2g_pytest.log 2>&1; echo pytest_rc=$?; tail -4 gpurun_out/r2g_pytest.log
for spc in 0 148; do SOSIE_GRP_SPC=$spc timeout 300 python tools/bench_kernels.py --tma 0 2>&1 | tail -1; done
timeout 300 python tools/bench_kernels.py --tma 0 --slabs 1460 2>&1 | tail -1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_apply' -c 5 --csv --log-file gpurun_out/r2g_launches.csv python tools/bench_kernels.py --tma 0 --reps 1 > gpurun_out/r2g_l.log 2>&1
grep -E "k_apply" gpurun_out/r2g_launches.csv | awk -F'","' '{print substr($5,1,30), $(NF-1), $NF}' | tail -5
