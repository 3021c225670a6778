for seed in 7 8 9 10; do
timeout 1500 python tools/fuzz_parity.py 600 $seed > gpurun_out/fuzz$seed.log 2>&1; echo rc=$?
grep -E "FAIL|ERROR|fuzz:" gpurun_out/fuzz$seed.log | cut -c1-300 | head -12
done
