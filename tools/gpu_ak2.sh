# Akima with ZF8 built in place: parity (GPU suite + akima-heavy fuzz) and the three timing shapes
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2r_pytest.log 2>&1; echo pytest_rc=$?; tail -4 gpurun_out/r2r_pytest.log
for seed in 201 202 203; do
timeout 600 python tools/fuzz_parity.py 2500 $seed > gpurun_out/fuzz_r2_$seed.log 2>&1; echo rc=$?
grep -E "FAIL|ERROR|fuzz:" gpurun_out/fuzz_r2_$seed.log | cut -c1-300 | head -6
done
timeout 300 python tools/bench_akima.py 2>&1 | tail -3
timeout 300 python tools/bench_akima_E.py 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
