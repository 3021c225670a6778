set -x
nvidia-smi -L
free -g | head -2; nproc
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2a_pytest.log 2>&1; echo pytest_rc=$?; tail -15 gpurun_out/r2a_pytest.log
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu --quick > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; echo bench_rc=$?
tail -c 1500 gpurun_out/r2a_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2a_bench.json'))
print('value',d['value'],'ms',d['ms_per_step'],'e2e_ms',d['e2e']['ms_per_step'],'roof',d['roofline'].get('traffic'),d['roofline'].get('traffic_note'))
for k in ('D','C'):
    print(k, json.dumps(d['others'][k].get('call_patterns'),indent=1))
PY
