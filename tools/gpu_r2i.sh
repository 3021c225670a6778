set -x
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2i_pytest.log 2>&1; echo pytest_rc=$?; tail -6 gpurun_out/r2i_pytest.log
timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu --no-others > gpurun_out/r2i_bench.json 2> gpurun_out/r2i_bench.err; echo bench_rc=$?
tail -c 600 gpurun_out/r2i_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2i_bench.json'))
print('value',d['value'],'ms',d['ms_per_step'],'map_ms',d['roofline']['kernel_ms'],'e2e_ms',d['e2e']['ms_per_step'],'launches',d['gpu_launches'],'equal',d['host_vs_device_path_equal'])
PY
