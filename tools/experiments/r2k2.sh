#!/bin/bash
# on the GPU box: variant sweep (per-kernel times), gzip parity subset for the CRC kernel, ncu of resolve (default build) and CRC
mkdir -p gpurun_out
tag=${1:-r2k2}
timeout -s KILL 600 python -m pytest tests -m gpu -x -q -k "gzip or config4 or crc or edge" > gpurun_out/${tag}_pytest.log 2>&1; tail -3 gpurun_out/${tag}_pytest.log
bash tools/sweep.sh --no-also > gpurun_out/${tag}_sweep.log 2>&1
grep -E "^==|^\[tbz\]" gpurun_out/${tag}_sweep.log
TBZ_KTIME=1 timeout -s KILL 200 python bench.py --workload gzip1m --members 2048 --steps 3 --warmup 2 --e2e-steps 1 --cpu-sample 1 --no-also 2>&1 | grep -E "^\[tbz\]" | tail -2
timeout -s KILL 400 ncu --set full --clock-control none --import-source on -k regex:k_inflate_resolve -s 3 -c 1 -o gpurun_out/${tag} python bench.py --steps 2 --warmup 3 --e2e-steps 1 --cpu-sample 16 --no-also > gpurun_out/${tag}_b.log 2>&1
timeout -s KILL 400 ncu --set full --clock-control none --import-source on -k regex:k_member_crc -s 2 -c 1 -o gpurun_out/${tag}_crc python bench.py --workload gzip1m --members 2048 --steps 2 --warmup 2 --e2e-steps 1 --cpu-sample 1 --no-also > gpurun_out/${tag}_crc_b.log 2>&1
