for seed in 1 2 3; do
timeout 1200 python tools/fuzz_parity.py 400 $seed > gpurun_out/fuzz$seed.log 2>&1; echo rc=$?
grep -E "FAIL|ERROR|fuzz:" gpurun_out/fuzz$seed.log | cut -c1-400 | head -20
done
