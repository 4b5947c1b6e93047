#!/bin/bash
# on the GPU box: after the stricter distance-code rule and the BFINAL = 0 filter: parity, dropped candidates, stage times
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests -m gpu -x -q -k "split" > gpurun_out/r2fp2_pytest.log 2>&1; tail -2 gpurun_out/r2fp2_pytest.log
for kb in 96 160 224 256; do
  echo "== chunk ${kb} KiB"
  TBZ_SPLIT_CHUNK_KB=$kb TBZ_KTIME=1 timeout -s KILL 300 python bench.py --workload gzip1g --steps 2 --warmup 1 --e2e-steps 1 --cpu-sample 1 --no-also 2> gpurun_out/r2fp2_err.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('  %.1f GB/s  %.2f ms  %s'%(d['value'], d['ms_per_step'], d['verification']['ok']))"
  grep "tbz split" gpurun_out/r2fp2_err.log | tail -14 | grep -v "slowest decode"
done 2>&1 | tee gpurun_out/r2fp2.log
