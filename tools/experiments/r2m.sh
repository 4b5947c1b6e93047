#!/bin/bash
# on the GPU box: e2e pipeline with small leading parts (TBZ_PIPE_FIRST = divisor of the first two parts; 0 = equal parts), split decode stages
mkdir -p gpurun_out
tag=${1:-r2m}
timeout -s KILL 600 python -m pytest tests -m gpu -x -q -k "split or gzip or batch" > gpurun_out/${tag}_pytest.log 2>&1; tail -2 gpurun_out/${tag}_pytest.log
for F in 0 16 32 8 16 0; do
  echo "TBZ_PIPE_FIRST=$F"
  TBZ_PIPE_FIRST=$F timeout -s KILL 300 python bench.py --no-also --steps 5 --warmup 3 --cpu-sample 16 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('  e2e %.2f GB/s %.3f ms (ceiling %.1f)  device %.1f GB/s'%(d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e']['ceiling_gbs'], d['value']))"
done 2>&1 | tee gpurun_out/${tag}_e2e.log
TBZ_KTIME=1 timeout -s KILL 300 python bench.py --workload gzip1g --steps 2 --warmup 2 --e2e-steps 1 --cpu-sample 1 --no-also 2>&1 | grep "tbz split" | tail -7 | tee gpurun_out/${tag}_split.log
