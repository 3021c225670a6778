set -x
python bench.py --steps 1 --warmup 3 --no-cpu --no-others --scale 0.5 > gpurun_out/plainM.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'k_map_easy' -s 4 -c 1 -o gpurun_out/prof_map_s2 -f python bench.py --steps 1 --warmup 3 --no-cpu --no-others --scale 0.5 > gpurun_out/ncuM.log 2>&1
tail -2 gpurun_out/ncuM.log
