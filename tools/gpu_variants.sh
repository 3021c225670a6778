for v in "" _variants/lib_fb6.so _variants/lib_fb8.so; do
  if [ -n "$v" ]; then export SOSIE_B200_LIB=$PWD/$v; fi
  python bench.py --steps 5 --warmup 3 --no-cpu --no-others --no-e2e 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print('$v', 'ms', round(d['ms_per_step'],3), 'map_ms', round(d['roofline']['kernel_ms'],3), 'apply', round(d['kernels'][0]['ms'],3))"
done
