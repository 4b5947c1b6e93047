#!/bin/bash
# on the GPU box: split decode stage times for several chunk sizes (why 224 and 320 KiB are fast and 236 / 256 / 384 are not)
mkdir -p gpurun_out
for kb in 224 256 192 288; do
  echo "== chunk ${kb} KiB"
  TBZ_SPLIT_CHUNK_KB=$kb TBZ_KTIME=1 timeout -s KILL 300 python bench.py --workload gzip1g --steps 2 --warmup 2 --e2e-steps 1 --cpu-sample 1 --no-also 2> gpurun_out/r2w2_err.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('  %.1f GB/s  %.2f ms'%(d['value'], d['ms_per_step']))"
  grep "tbz split" gpurun_out/r2w2_err.log | tail -7
done 2>&1 | tee gpurun_out/r2w2_stages.log
