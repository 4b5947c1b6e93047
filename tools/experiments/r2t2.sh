#!/bin/bash
# on the GPU box: split decode after the Kraft early exit in the block-start validation and the 16 Ki-symbol ring: parity, stages per variant, chunk sweep
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests -m gpu -x -q -k "split" > gpurun_out/r2t2_pytest.log 2>&1; tail -2 gpurun_out/r2t2_pytest.log
for so in 3bz_b200/libthreebz_cuda.so 3bz_b200/var_ring32k.so 3bz_b200/var_ring8k.so; do
  echo "== $so"
  TBZ_LIB=$PWD/$so TBZ_KTIME=1 timeout -s KILL 300 python bench.py --workload gzip1g --steps 2 --warmup 2 --e2e-steps 1 --cpu-sample 1 --no-also 2>&1 | grep "tbz split" | tail -7
done 2>&1 | tee gpurun_out/r2t2_stages.log
KBS="224 160 128 96" bash tools/split_sweep.sh 2>&1 | tee gpurun_out/r2t2_split.log
