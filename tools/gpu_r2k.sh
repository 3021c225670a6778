set -x
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2k_pytest.log 2>&1; echo pytest_rc=$?; tail -4 gpurun_out/r2k_pytest.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu --no-others > gpurun_out/r2k_bench.json 2> gpurun_out/r2k_bench.err; echo rc=$?
python - <<PY
import json
t=open('gpurun_out/r2k_bench.json').read(); d=json.loads(t[t.index('{"metric'):].split('\n')[0])
print('N',d['n_gpus'],'value %.3e'%d['value'],'ms',round(d['ms_per_step'],3),'map',d['roofline']['kernel_ms'],'e2e_ms',round(d['e2e']['ms_per_step'],2),'equal',d['host_vs_device_path_equal'],'launches',d['gpu_launches'])
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/r2k_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu --no-others --no-e2e > gpurun_out/r2k_l.log 2>&1
