for mode in on off; do
  if [ $mode = off ]; then export SOSIE_BENCH_NO_EXCHANGE=1; else unset SOSIE_BENCH_NO_EXCHANGE; fi
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 3 --warmup 3 --no-cpu --no-others > gpurun_out/bench_n8_$mode.json 2> gpurun_out/bench_n8_$mode.err; echo rc=$?
  tail -3 gpurun_out/bench_n8_$mode.err
  python -c "
import json; d=json.loads(open('gpurun_out/bench_n8_$mode.json').read().strip().split('\n')[-1])
print('$mode value',d['value'],'ms',d['ms_per_step'],'e2e_ms',d['e2e']['ms_per_step'],'nomap',d['e2e']['without_mapping_download']['ms_per_step'],'xchg',d['e2e'].get('prelude_exchange'),'h2d',d['e2e']['h2d_bytes_per_step'],'equal',d['host_vs_device_path_equal'])"
done
