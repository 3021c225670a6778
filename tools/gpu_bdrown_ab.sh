# on-chip BDROWN on C: timing + the tests that cover BDROWN / SMOOTHER
timeout 300 python tools/bench_kernels.py --what bdrownC 2>&1 | tail -1
timeout 900 python -m pytest tests -m gpu -q -x -k "bdrown or smoother or drown or fuzz or 3d_job or golden" 2>&1 | tail -2
