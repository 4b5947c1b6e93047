#!/bin/bash
# round 2, first GPU pass: parity, per-kernel times of the variants, ncu of the new phase two
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2b_smi.txt 2>&1
timeout -s KILL 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2b_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2b_pytest.log
tail -5 gpurun_out/r2b_pytest.log
bash tools/sweep.sh > gpurun_out/r2b_sweep.log 2>&1
cat gpurun_out/r2b_sweep.log
timeout -s KILL 300 python bench.py --steps 20 --warmup 5 > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err; tail -c 1500 gpurun_out/r2b_bench.json
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:k_inflate_ -s 6 -c 2 -o gpurun_out/r2b python bench.py --steps 2 --warmup 3 --e2e-steps 1 --cpu-sample 16 > gpurun_out/r2b_ncu.log 2>&1
ls -la gpurun_out | tail -5
