set -x
timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -k "async or exchange" > gpurun_out/r2j_pytest.log 2>&1; echo pytest_rc=$?; tail -6 gpurun_out/r2j_pytest.log
for n in 8 2; do
SOSIE_BENCH_DEBUG=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 20 --warmup 3 > gpurun_out/r2_bench_n$n.json 2> gpurun_out/r2_bench_n$n.err; echo rc=$?
grep "rank 0" gpurun_out/r2_bench_n$n.err | head -2
python - <<PY
import json
t=open('gpurun_out/r2_bench_n$n.json').read(); d=json.loads(t[t.index('{"metric'):].split('\n')[0])
print('N',d['n_gpus'],'value %.3e'%d['value'],'ms',round(d['ms_per_step'],3),'e2e_ms',round(d['e2e']['ms_per_step'],2),'equal',d['host_vs_device_path_equal'])
PY
done
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu --no-others > gpurun_out/r2j_bench.json 2> gpurun_out/r2j_bench.err; echo rc=$?
python - <<PY
import json
t=open('gpurun_out/r2j_bench.json').read(); d=json.loads(t[t.index('{"metric'):].split('\n')[0])
print('N',d['n_gpus'],'value %.3e'%d['value'],'ms',round(d['ms_per_step'],3),'e2e_ms',round(d['e2e']['ms_per_step'],2),'equal',d['host_vs_device_path_equal'])
PY
