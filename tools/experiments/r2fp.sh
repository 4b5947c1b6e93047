#!/bin/bash
# on the GPU box: which candidates of the block-start search are false positives (TBZ_KTIME prints their header fields)
mkdir -p gpurun_out
for kb in 64 96 160 256; do
  echo "== chunk ${kb} KiB"
  TBZ_SPLIT_CHUNK_KB=$kb TBZ_KTIME=1 timeout -s KILL 300 python bench.py --workload gzip1g --steps 1 --warmup 1 --e2e-steps 1 --cpu-sample 1 --no-also 2>&1 >/dev/null | grep -E "dropped|chunks of" | sort | uniq -c | sort -rn | head -14
done 2>&1 | tee gpurun_out/r2fp.log
