#!/bin/bash
# on the GPU box: the block-start search with lane-parallel validation: parity, the same chunks found, stage times
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests -m gpu -x -q -k "split" > gpurun_out/r2v2_pytest.log 2>&1; tail -2 gpurun_out/r2v2_pytest.log
for kb in 160 0; do
  echo "== chunk ${kb} KiB (0 = default)"
  if [ $kb = 0 ]; then unset TBZ_SPLIT_CHUNK_KB; else export TBZ_SPLIT_CHUNK_KB=$kb; fi
  TBZ_KTIME=1 timeout -s KILL 300 python bench.py --workload gzip1g --steps 2 --warmup 2 --e2e-steps 1 --cpu-sample 1 --no-also 2> gpurun_out/r2v2_err.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('  %.1f GB/s  %.2f ms  verified %s'%(d['value'], d['ms_per_step'], d['verification']))"
  grep "tbz split" gpurun_out/r2v2_err.log | tail -7
done 2>&1 | tee gpurun_out/r2v2_stages.log
