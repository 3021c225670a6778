set -x
B="python bench.py --steps 2 --warmup 3 --no-cpu --no-others --no-e2e"
$B > gpurun_out/plain_E.log 2>&1 || { tail -5 gpurun_out/plain_E.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_E_r01e.csv $B > gpurun_out/ncu_l.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_map_easy|k_apply1|k_pack' -s 9 -c 3 -o gpurun_out/prof_E_r01e -f $B > gpurun_out/ncu_f.log 2>&1
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo rc=$?
tail -c 300 gpurun_out/bench_full.err
