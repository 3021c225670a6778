timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "track" 2>&1 | tail -15
