N=$1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo rc=$?
tail -3 gpurun_out/bench_n$N.err
python -c "
import json; d=json.loads(open('gpurun_out/bench_n$N.json').read().strip().split('\n')[-1])
print('N=$N value',d['value'],'ms',d['ms_per_step'],'map_ms',d['roofline']['kernel_ms'],'e2e_ms',d['e2e']['ms_per_step'],'e2e',d['e2e']['value'],'equal',d['host_vs_device_path_equal'])"
