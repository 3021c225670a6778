B="python bench.py --steps 2 --warmup 3 --no-cpu --no-others --no-e2e"
$B > gpurun_out/plain_E.log 2>&1 || { tail -5 gpurun_out/plain_E.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:'k_map_easy' -s 3 -c 1 -o gpurun_out/prof_map_p3 -f $B > gpurun_out/ncu_f.log 2>&1
tail -2 gpurun_out/ncu_f.log
