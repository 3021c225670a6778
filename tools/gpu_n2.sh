python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo rc=$?
tail -5 gpurun_out/bench_n2.err
python -c "
import json; d=json.loads(open('gpurun_out/bench_n2.json').read().strip().split('\n')[-1])
print('value',d['value'],'ms',d['ms_per_step'],'map_ms',d['roofline']['kernel_ms'],'e2e',d['e2e'],'equal',d['host_vs_device_path_equal'])"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 | tail -1 | cut -c1-300
