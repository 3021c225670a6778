for seed in 4 5 6; do
timeout 1200 python tools/fuzz_parity.py 500 $seed > gpurun_out/fuzz$seed.log 2>&1; echo rc=$?
grep -E "FAIL|ERROR|fuzz:" gpurun_out/fuzz$seed.log | cut -c1-400 | head -20
done
timeout 900 python -m pytest tests/test_gpu_fuzz.py -x -q 2>&1 | tail -3
