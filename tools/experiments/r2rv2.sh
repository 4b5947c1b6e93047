#!/bin/bash
# on the GPU box: symbolic resolve with 512 (default) / 768 / 1024 threads per CTA; parity of the split path
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests -m gpu -x -q -k "split" > gpurun_out/r2rv2_pytest.log 2>&1; tail -2 gpurun_out/r2rv2_pytest.log
for so in 3bz_b200/libthreebz_cuda.so 3bz_b200/var_res768.so 3bz_b200/var_res1024.so; do
  echo "== $so"
  TBZ_LIB=$PWD/$so TBZ_KTIME=1 timeout -s KILL 300 python bench.py --workload gzip1g --steps 2 --warmup 1 --e2e-steps 1 --cpu-sample 1 --no-also 2> gpurun_out/r2rv2_err.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('  %.1f GB/s  %.2f ms  %s'%(d['value'], d['ms_per_step'], d['verification']['ok']))"
  grep "tbz split" gpurun_out/r2rv2_err.log | tail -8 | grep -E "resolve|decode  "
done 2>&1 | tee gpurun_out/r2rv2.log
