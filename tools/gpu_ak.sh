timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_fuzz.py -x -q -k "akima or Akima or fuzz" 2>&1 | tail -3
python tools/bench_akima_E.py 2>&1 | tail -1 | cut -c1-200
python tools/bench_akima.py 16 2>&1 | tail -1
python tools/bench_akima_E.py 2>&1 | tail -1 | cut -c1-200
