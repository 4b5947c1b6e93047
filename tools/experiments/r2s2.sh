#!/bin/bash
# on the GPU box: ncu --set full of the split decode's three heavy kernels after the third session's changes (config 3, one launch each)
mkdir -p gpurun_out
timeout -s KILL 250 ncu --set full --clock-control none --import-source on -k regex:"k_split_find|k_split_decode|k_split_resolve" -c 3 -o gpurun_out/r2s2_split python bench.py --workload gzip1g --steps 1 --warmup 0 --e2e-steps 1 --cpu-sample 1 --no-also > gpurun_out/r2s2_b.log 2>&1
ls -la gpurun_out/r2s2_split.ncu-rep
