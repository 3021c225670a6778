timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -x -q -k "akima or Akima" 2>&1 | tail -4
python tools/bench_akima_E.py 2>&1 | tail -1
python tools/bench_akima.py 16 2>&1 | tail -1
