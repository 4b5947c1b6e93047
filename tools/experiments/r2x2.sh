#!/bin/bash
# on the GPU box: block-start search with two-stage compaction: parity; slowest chunks of the split decode at a fast and a slow chunk size
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests -m gpu -x -q -k "split" > gpurun_out/r2x2_pytest.log 2>&1; tail -2 gpurun_out/r2x2_pytest.log
for kb in 160 224 256; do
  echo "== chunk ${kb} KiB"
  TBZ_SPLIT_CHUNK_KB=$kb TBZ_KTIME=1 timeout -s KILL 300 python bench.py --workload gzip1g --steps 2 --warmup 2 --e2e-steps 1 --cpu-sample 1 --no-also 2> gpurun_out/r2x2_err.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('  %.1f GB/s  %.2f ms'%(d['value'], d['ms_per_step']))"
  grep "tbz split" gpurun_out/r2x2_err.log | tail -12
done 2>&1 | tee gpurun_out/r2x2_stages.log
