timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python tools/bench_mapC.py C 2>&1 | tail -4
python tools/bench_mapC.py B 2>&1 | tail -4
timeout 900 python tools/fuzz_parity.py 600 21 2>&1 | tail -1
timeout 900 python tools/fuzz_parity.py 600 22 2>&1 | tail -1
SOSIE_TWISTED_NO_PRUNE=1 python tools/bench_mapC.py C 2>&1 | tail -2
