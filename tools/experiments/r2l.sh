#!/bin/bash
# on the GPU box: parity subset, resolve-occupancy variants, chunk-size sweep of the split decode with the CTA-per-span CRC
mkdir -p gpurun_out
tag=${1:-r2l}
timeout -s KILL 900 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; tail -3 gpurun_out/${tag}_pytest.log
bash tools/sweep.sh --no-also > gpurun_out/${tag}_sweep.log 2>&1
grep -E "^==|^\[tbz\]" gpurun_out/${tag}_sweep.log
KBS="160 112 80 56" bash tools/split_sweep.sh 2>&1 | tee gpurun_out/${tag}_split.log
TBZ_KTIME=1 timeout -s KILL 300 python bench.py --workload gzip1g --steps 2 --warmup 2 --e2e-steps 1 --cpu-sample 1 --no-also 2>&1 | grep "tbz split" | tail -7
