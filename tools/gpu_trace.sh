n=$1
SOSIE_PIPE_TRACE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 3 --warmup 3 --no-others --no-cpu > gpurun_out/trace_n$n.json 2> gpurun_out/trace_n$n.err
grep "sosie pipe" gpurun_out/trace_n$n.err | tail -60
