set -x
timeout 600 python -m pytest tests/test_gpu_tma.py tests/test_gpu_queue.py tests/test_gpu_fullsize.py -q -x > gpurun_out/r2b_pytest.log 2>&1; echo pytest_rc=$?; tail -25 gpurun_out/r2b_pytest.log
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2b_pytest_all.log 2>&1; echo pytest_all_rc=$?; tail -5 gpurun_out/r2b_pytest_all.log
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu --quick > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err; echo bench_rc=$?
tail -c 800 gpurun_out/r2b_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2b_bench.json'))
print('value',d['value'],'ms',d['ms_per_step'],'e2e_ms',d['e2e']['ms_per_step'])
print('D', {k:v for k,v in d['others']['D'].items() if k!='call_patterns'})
for k in ('D','C'):
    print(k, json.dumps(d['others'][k].get('call_patterns')))
PY
(timeout 300 python tools/bench_kernels.py --tma 0; timeout 300 python tools/bench_kernels.py --tma 4) > gpurun_out/r2b_kern.log 2>&1; tail -20 gpurun_out/r2b_kern.log
