timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "exchange or pipelined" 2>&1 | tail -15
