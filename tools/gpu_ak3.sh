SOSIE_B200_LIB=$PWD/_variants/lib_old.so python tools/bench_akima.py 16 2>&1 | tail -1 | cut -c1-120
python tools/bench_akima.py 16 2>&1 | tail -1 | cut -c1-120
SOSIE_B200_LIB=$PWD/_variants/lib_old.so python tools/bench_akima.py 16 2>&1 | tail -1 | cut -c1-120
python tools/bench_akima.py 16 2>&1 | tail -1 | cut -c1-120
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_akC_new.csv python tools/bench_akima.py 16 > /dev/null 2>&1
python tools/ncu_summary.py list gpurun_out/launches_akC_new.csv gpurun_out/launches_akC_new.md "akC new"
head -30 gpurun_out/launches_akC_new.md
