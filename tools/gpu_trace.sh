SOSIE_PIPE_TRACE=1 python bench.py --steps 3 --warmup 3 --no-cpu --no-others > gpurun_out/bench_t.json 2> gpurun_out/bench_t.err; echo rc=$?
tail -45 gpurun_out/bench_t.err
python -c "
import json; d=json.load(open('gpurun_out/bench_t.json'))
print('value',d['value'],'ms',d['ms_per_step'],'map_ms',d['roofline']['kernel_ms'],'e2e',d['e2e'])"
nvidia-smi --query-gpu=pcie.link.gen.current,pcie.link.width.current,pcie.link.gen.max --format=csv
