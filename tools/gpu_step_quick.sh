timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_golden.py tests/test_gpu_tma.py tests/test_gpu_fullsize.py -m gpu -q -x 2>&1 | tail -2
for rep in 1 2; do
python bench.py --steps 8 --warmup 3 --no-cpu --no-others --no-e2e 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('ms_per_step', round(d['ms_per_step'], 4), 'map_kernel_ms', round(d['roofline']['kernel_ms'], 4), [ (k.get('name'), round(k.get('ms', 0), 4)) for k in d.get('kernels', []) ][:4])
"
done
