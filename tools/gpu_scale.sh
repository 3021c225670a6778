for n in 1 2 4 8; do
if [ $n = 1 ]; then
( time timeout 1500 python bench.py --gpus 1 --steps 20 --warmup 3 > gpurun_out/scale_r02_n1.json 2> gpurun_out/scale_r02_n1.err ) 2>&1 | grep real
else
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 20 --warmup 3 > gpurun_out/scale_r02_n$n.json 2> gpurun_out/scale_r02_n$n.err ) 2>&1 | grep real
fi
python - <<PY
import json
t=open('gpurun_out/scale_r02_n$n.json').read(); d=json.loads(t[t.index('{"metric'):].split('\n')[0])
o=d.get('others',{})
print('N',d['n_gpus'],'value %.4e'%d['value'],'ms %.3f'%d['ms_per_step'],'e2e_ms %.2f'%d['e2e']['ms_per_step'],'clk',d['clocks'].get('sm_mhz'),d['clocks'].get('reasons'),
      'D',{k:round(v,3) for k,v in o.get('D',{}).get('sharded',{}).items() if isinstance(v,float)},
      'C',{k:round(v,3) for k,v in o.get('C',{}).get('sharded',{}).items() if isinstance(v,float)})
PY
done
( time timeout 600 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > gpurun_out/scale_r02_ref.json 2> gpurun_out/scale_r02_ref.err ) 2>&1 | grep real
cat gpurun_out/scale_r02_ref.json | cut -c1-600
