# E step with experiment builds of one csrc file (tools/build_variant.py): ms_per_step and the map kernel's own time
for v in "" $VARIANTS; do
  if [ -n "$v" ]; then export SOSIE_B200_LIB=$PWD/variants/libsosie_$v.so; fi
  for rep in 1 2; do
  python bench.py --steps 8 --warmup 3 --no-cpu --no-others --no-e2e 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('variant=${v:-default}', 'ms_per_step', round(d['ms_per_step'], 4), 'map_kernel_ms', round(d['roofline']['kernel_ms'], 4), 'sm_mhz', d['clocks']['sm_mhz'])
"
  done
done
if [ -n "$CHECK" ]; then export SOSIE_B200_LIB=$PWD/variants/libsosie_$CHECK.so; timeout 600 python -m pytest tests/test_golden.py tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -3; fi
