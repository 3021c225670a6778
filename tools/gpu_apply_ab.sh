# batched apply on D: timing at 292 / 584 / 1460 slabs + the parity tests that cover the batched kernels
for S in 292 584 1460; do timeout 300 python tools/bench_kernels.py --tma 0 --slabs $S 2>&1 | tail -1; done
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_queue.py tests/test_gpu_tma.py tests/test_golden.py -m gpu -q -x 2>&1 | tail -3
