for v in "" $VARIANTS "" $VARIANTS; do
  if [ -n "$v" ]; then export SOSIE_B200_LIB=$PWD/variants/libsosie_$v.so; else unset SOSIE_B200_LIB; fi
  echo -n "variant=${v:-default} "; timeout 300 python tools/bench_kernels.py --what bdrownC --reps 21 2>&1 | tail -1
done
