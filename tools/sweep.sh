#!/bin/bash
# on the GPU box: per-kernel times of the default build and every 3bz_b200/var_*.so on the bench workload
for so in 3bz_b200/libthreebz_cuda.so 3bz_b200/var_*.so; do
  [ -e "$so" ] || continue
  echo "== $so"
  TBZ_LIB=$PWD/$so TBZ_KTIME=1 timeout -s KILL 120 python bench.py --steps 3 --warmup 3 --e2e-steps 1 --cpu-sample 16 "$@" 2>&1 | grep -E "^\[tbz\]|ms_per_step" | sed -n '4p;$p' | cut -c1-160
done
