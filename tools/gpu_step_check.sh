timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/r2_step_pytest.log 2>&1; echo pytest_rc=$?; tail -2 gpurun_out/r2_step_pytest.log
for rep in 1 2; do
python bench.py --steps 8 --warmup 3 --no-cpu --no-others --no-e2e 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('ms_per_step', round(d['ms_per_step'], 4), 'map_kernel_ms', round(d['roofline']['kernel_ms'], 4), [ (k.get('name'), round(k.get('ms', 0), 4)) for k in d.get('kernels', []) ][:4])
"
done
timeout 300 python tools/fuzz_parity.py 2500 501 2>&1 | grep -E "FAIL|fuzz:" | head -5
