set -x
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_map_easy' -s 3 -c 1 -o gpurun_out/prof_E_r02a -f python bench.py --steps 2 --warmup 3 --no-cpu --no-others --no-e2e > gpurun_out/r2m_ncu.log 2>&1
tail -2 gpurun_out/r2m_ncu.log
