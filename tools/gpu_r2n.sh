for mb in 6 10 12; do
SOSIE_B200_LIB=$PWD/_variants/mb$mb/libsosie_b200.so timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu --no-others --no-e2e > gpurun_out/r2n_$mb.json 2> gpurun_out/r2n_$mb.err
python - <<PY
import json
t=open('gpurun_out/r2n_$mb.json').read(); d=json.loads(t[t.index('{"metric'):].split('\n')[0])
print('MB $mb','ms',round(d['ms_per_step'],3),'map',round(d['roofline']['kernel_ms'],3))
PY
done
