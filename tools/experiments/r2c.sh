#!/bin/bash
# round 2: parity, bench and ncu (--set full, both hot kernels) of the working tree
tag=${1:-r2c}
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${tag}_pytest.log
tail -5 gpurun_out/${tag}_pytest.log
TBZ_KTIME=1 timeout -s KILL 300 python bench.py --steps 20 --warmup 5 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; tail -c 1500 gpurun_out/${tag}_bench.json; grep "^\[tbz\]" gpurun_out/${tag}_bench.err | sed -n '6p'
timeout -s KILL 400 ncu --set full --clock-control none --import-source on -k regex:k_inflate_ -s 6 -c 2 -o gpurun_out/${tag} python bench.py --steps 2 --warmup 3 --e2e-steps 1 --cpu-sample 16 > gpurun_out/${tag}_ncu.log 2>&1
ls -la gpurun_out | tail -5
