python -m pytest tests -m gpu -x -q -k "bdrown or smoother or drown or config_C or config_A" > gpurun_out/pytest_drown.log 2>&1; tail -5 gpurun_out/pytest_drown.log
python tools/bench_drown_big.py > gpurun_out/drown_big.json 2> gpurun_out/drown_big.err; echo rc=$?; tail -3 gpurun_out/drown_big.err; cat gpurun_out/drown_big.json
