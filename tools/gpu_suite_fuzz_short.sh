# full GPU suite + three fuzz seeds + the E step (used after kernel changes)
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2s_pytest.log 2>&1; echo pytest_rc=$?; tail -3 gpurun_out/r2s_pytest.log
for seed in ${SEEDS:-301 302 303}; do
timeout 600 python tools/fuzz_parity.py 2500 $seed > gpurun_out/fuzz_r2_$seed.log 2>&1; echo rc=$?
grep -E "FAIL|ERROR|fuzz:" gpurun_out/fuzz_r2_$seed.log | cut -c1-300 | head -6
done
python bench.py --steps 8 --warmup 3 --no-cpu --no-others --no-e2e 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('ms_per_step', round(d['ms_per_step'], 4), 'map_kernel_ms', round(d['roofline']['kernel_ms'], 4))
"
