set -x
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
python bench.py --steps 5 --warmup 3 --no-cpu --no-others > gpurun_out/bench_q.json 2> gpurun_out/bench_q.err; echo rc=$?
tail -c 600 gpurun_out/bench_q.err
python -c "
import json; d=json.load(open('gpurun_out/bench_q.json'))
print('value',d['value'],'ms',d['ms_per_step'],'map_ms',d['roofline']['kernel_ms'],'e2e_ms',d['e2e']['ms_per_step'],'launches',d['gpu_launches'],'equal',d['host_vs_device_path_equal'])"
