set -x
for n in 8 4 2; do
SOSIE_BENCH_DEBUG=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 20 --warmup 3 > gpurun_out/r2_bench_n$n.json 2> gpurun_out/r2_bench_n$n.err; echo rc=$?
grep "rank" gpurun_out/r2_bench_n$n.err | head -8
python - <<PY
import json
d=json.load(open('gpurun_out/r2_bench_n$n.json'))
print('N',d['n_gpus'],'value',d['value'],'ms',d['ms_per_step'],'e2e_ms',d['e2e']['ms_per_step'],'xchg',d['e2e'].get('prelude_exchange'),'equal',d['host_vs_device_path_equal'])
PY
done
