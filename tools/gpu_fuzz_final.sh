for seed in ${SEEDS:-401 402 403 404 405 406 407 408}; do
timeout 600 python tools/fuzz_parity.py 2500 $seed > gpurun_out/fuzz_r2_$seed.log 2>&1; echo rc=$?
grep -E "FAIL|ERROR|fuzz:" gpurun_out/fuzz_r2_$seed.log | cut -c1-300 | head -6
done
