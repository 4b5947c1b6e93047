#!/bin/bash
# on the GPU box: sub-chunk size of the split decode's phase one (one warp per chunk, <= 11 warps per SM: L1 is not contended here)
mkdir -p gpurun_out
for so in 3bz_b200/var_ds3000.so 3bz_b200/var_ds5000.so 3bz_b200/var_ds6000.so; do
  echo "== $so"
  TBZ_LIB=$PWD/$so TBZ_KTIME=1 timeout -s KILL 300 python bench.py --workload gzip1g --steps 2 --warmup 1 --e2e-steps 1 --cpu-sample 1 --no-also 2> gpurun_out/r2ds_err.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('  %.1f GB/s  %.2f ms  %s'%(d['value'], d['ms_per_step'], d['verification']['ok']))"
  grep "tbz split" gpurun_out/r2ds_err.log | tail -8 | grep -E "decode  |resolve"
done 2>&1 | tee gpurun_out/r2ds.log
