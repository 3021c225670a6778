for v in _variants/*/; do
  SOSIE_B200_LIB=$PWD/${v}libsosie_b200.so python bench.py --steps 5 --warmup 3 --no-cpu --no-others > gpurun_out/bench_v.json 2>gpurun_out/bench_v.err || tail -3 gpurun_out/bench_v.err
  python -c "
import json; d=json.load(open('gpurun_out/bench_v.json'))
print('$v','ms',round(d['ms_per_step'],3),'map_ms',round(d['roofline']['kernel_ms'],3),'e2e_ms',round(d['e2e']['ms_per_step'],2))"
done
python bench.py --steps 5 --warmup 3 --no-cpu --no-others > gpurun_out/bench_v.json 2>gpurun_out/bench_v.err
python -c "
import json; d=json.load(open('gpurun_out/bench_v.json'))
print('default','ms',round(d['ms_per_step'],3),'map_ms',round(d['roofline']['kernel_ms'],3),'e2e_ms',round(d['e2e']['ms_per_step'],2))"
ncu --set full --clock-control none --import-source on -k regex:'k_map_easy' -s 4 -c 1 -o gpurun_out/prof_map_s3 -f python bench.py --steps 1 --warmup 3 --no-cpu --no-others --scale 0.5 > gpurun_out/ncuM.log 2>&1
tail -1 gpurun_out/ncuM.log
