set -x
timeout 600 python -m pytest tests/test_gpu_tma.py tests/test_gpu_queue.py tests/test_gpu_fullsize.py -q -x > gpurun_out/r2d_pytest.log 2>&1; echo pytest_rc=$?; tail -4 gpurun_out/r2d_pytest.log
for spc in 0 32 64 148 292; do SOSIE_GRP_SPC=$spc timeout 300 python tools/bench_kernels.py --tma 0 2>&1 | tail -1; done
timeout 300 python tools/bench_kernels.py --tma 0 --slabs 1460 2>&1 | tail -1
timeout 300 python tools/bench_kernels.py --tma 4 2>&1 | tail -1
