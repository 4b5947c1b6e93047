#!/bin/bash
# usage (on the GPU box): tools/prof1.sh <tag> <lib.so> [kernel regex]  -> gpurun_out/<tag>.ncu-rep (one --set full capture of a full batch)
tag=$1; lib=$2; k=${3:-k_inflate_resolve}
TBZ_LIB=$PWD/$lib timeout -s KILL 400 ncu --set full --clock-control none --import-source on -k regex:$k -s 3 -c 1 -o gpurun_out/$tag python bench.py --steps 2 --warmup 3 --e2e-steps 1 --cpu-sample 16 > gpurun_out/${tag}_b.log 2>&1
