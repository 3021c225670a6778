set -x
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo rc=$?
tail -c 3000 gpurun_out/bench_full.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2>> gpurun_out/bench_full.err; echo rc=$?
python bench.py --steps 1 --warmup 3 --no-cpu --no-others > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01.csv python bench.py --steps 1 --warmup 3 --no-cpu --no-others > gpurun_out/ncu1.log 2>&1
python bench.py --steps 1 --warmup 3 --no-cpu --no-others > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'k_map_easy|k_apply|k_pack' -s 6 -c 3 -o gpurun_out/prof_E_r01 python bench.py --steps 1 --warmup 3 --no-cpu --no-others > gpurun_out/ncu2.log 2>&1
ls -la gpurun_out
