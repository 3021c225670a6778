# the default bench on 2 GPUs of one box (the sharded mapping, the NCCL prelude exchange, the sharded D / C jobs)
n=${NG:-2}
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 20 --warmup 3 > gpurun_out/scale_r02_n$n.json 2> gpurun_out/scale_r02_n$n.err ) 2>&1 | grep real
tail -c 300 gpurun_out/scale_r02_n$n.err
python - <<PY
import json
t=open('gpurun_out/scale_r02_n$n.json').read(); d=json.loads(t[t.index('{"metric'):].split('\n')[0])
o=d.get('others',{})
print('N',d['n_gpus'],'value %.4e'%d['value'],'ms %.3f'%d['ms_per_step'],'e2e_ms %.2f'%d['e2e']['ms_per_step'],'clk',d['clocks'].get('sm_mhz'),d['clocks'].get('reasons'),
      'D',{k:round(v,3) for k,v in o.get('D',{}).get('sharded',{}).items() if isinstance(v,float)},
      'C',{k:round(v,3) for k,v in o.get('C',{}).get('sharded',{}).items() if isinstance(v,float)})
PY
python - <<PY
import json
t=open('gpurun_out/scale_r02_n$n.json').read(); d=json.loads(t[t.index('{"metric'):].split('\n')[0])
print('prelude_exchange', d.get('prelude_exchange'))
PY
grep -h "peer-memory\|p2p\|IPC" gpurun_out/scale_r02_n$n.err | head -5
