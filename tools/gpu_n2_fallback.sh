# N = 2: (1) the default bench (peer-memory exchange), (2) the same with a rank arriving late and a 200 ms bound: the fallback to NCCL
n=2
for mode in normal fail; do
  if [ $mode = fail ]; then export SOSIE_BENCH_P2P_TEST_FAIL=1 SOSIE_P2P_TIMEOUT_MS=200; fi
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 10 --warmup 3 --no-cpu --no-others --no-e2e > gpurun_out/n2_$mode.json 2> gpurun_out/n2_$mode.err; echo rc=$?
  grep -h "\[bench\]" gpurun_out/n2_$mode.err | head -4
  python - <<PY
import json
t=open('gpurun_out/n2_$mode.json').read(); d=json.loads(t[t.index('{"metric'):].split('\n')[0])
print('$mode', 'ms %.3f'%d['ms_per_step'], d.get('prelude_exchange'))
PY
done
