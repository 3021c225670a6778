python tools/bench_akima.py 16 2>&1 | tail -1
python tools/bench_akima.py 16 2>&1 | tail -1
SOSIE_B200_LIB=$PWD/_variants/lib_old.so python tools/bench_akima.py 16 2>&1 | tail -1
