# end-of-round evidence: new parity test, ncu captures (E step, D apply), then the default bench of both arms
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "index_nearest_ties" 2>&1 | tail -2
bash tools/gpu_profiles_r02.sh > gpurun_out/profiles_r02.log 2>&1; echo profiles_rc=$?
for S in 292 1460; do timeout 300 python tools/bench_kernels.py --tma 0 --slabs $S 2>&1 | tail -1; done
