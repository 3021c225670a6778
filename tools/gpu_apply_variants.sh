# batched apply on D with experiment builds (tools/build_variant.py)
for v in "" $VARIANTS; do
  if [ -n "$v" ]; then export SOSIE_B200_LIB=$PWD/variants/libsosie_$v.so; fi
  for S in 292 1460; do echo -n "variant=${v:-default} "; timeout 300 python tools/bench_kernels.py --tma 0 --slabs $S 2>&1 | tail -1; done
done
