FUZZ_ONLY=33,101,109,125 FUZZ_VERBOSE=1 timeout 600 python tools/fuzz_parity.py 240 1 > gpurun_out/fuzz_dbg.log 2>&1; echo rc=$?
