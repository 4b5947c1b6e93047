#!/bin/bash
# on the GPU box: symbolic resolve, threads per CTA around 512
mkdir -p gpurun_out
for so in 3bz_b200/var_res448.so 3bz_b200/var_res576.so 3bz_b200/var_res640.so 3bz_b200/var_res512t3.so; do
  echo "== $so"
  TBZ_LIB=$PWD/$so TBZ_KTIME=1 timeout -s KILL 300 python bench.py --workload gzip1g --steps 2 --warmup 1 --e2e-steps 1 --cpu-sample 1 --no-also 2> gpurun_out/r2rv3_err.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('  %.1f GB/s  %.2f ms  %s'%(d['value'], d['ms_per_step'], d['verification']['ok']))"
  grep "tbz split" gpurun_out/r2rv3_err.log | tail -8 | grep -E "resolve"
done 2>&1 | tee gpurun_out/r2rv3.log
