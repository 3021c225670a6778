set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -15
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_s2.json 2> gpurun_out/bench_s2.err; echo rc=$?
tail -c 1500 gpurun_out/bench_s2.err
