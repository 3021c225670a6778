set -x
n=2
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/r2l_bench_n$n.json 2> gpurun_out/r2l_bench_n$n.err; echo rc=$?
grep -v "^\[W\|Warning\|^\*\*\*\|OMP_NUM" gpurun_out/r2l_bench_n$n.err | tail -5
python - <<PY
import json
t=open('gpurun_out/r2l_bench_n$n.json').read(); d=json.loads(t[t.index('{"metric'):].split('\n')[0])
print('N',d['n_gpus'],'value %.3e'%d['value'],'ms',round(d['ms_per_step'],3),'e2e_ms',round(d['e2e']['ms_per_step'],2))
print(json.dumps(d.get('others'),indent=1))
PY
( time timeout 1200 python bench.py > gpurun_out/r2l_bench_full.json 2> gpurun_out/r2l_bench_full.err ) 2>&1 | tail -3; echo rc=$?
tail -c 400 gpurun_out/r2l_bench_full.err
