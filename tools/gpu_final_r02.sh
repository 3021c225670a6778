# end-of-round evidence: full GPU suite, smoke, ncu captures (E step, D apply), apply timings
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/r2_final_pytest.log 2>&1; echo pytest_rc=$?; tail -2 gpurun_out/r2_final_pytest.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
bash tools/gpu_profiles_r02.sh > gpurun_out/profiles_r02.log 2>&1; echo profiles_rc=$?
for S in 292 1460; do timeout 300 python tools/bench_kernels.py --tma 0 --slabs $S 2>&1 | tail -1; done
