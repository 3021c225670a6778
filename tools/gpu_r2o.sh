set -x
timeout 900 python -m pytest tests -m gpu -q -x -k "akima or golden or fuzz or smoke" > gpurun_out/r2o_pytest.log 2>&1; echo pytest_rc=$?; tail -6 gpurun_out/r2o_pytest.log
timeout 300 python tools/bench_akima_E.py 2>&1 | tail -3
timeout 300 python tools/bench_akima.py 2>&1 | tail -3
