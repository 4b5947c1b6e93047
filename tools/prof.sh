#!/bin/bash
# usage (on the GPU box): tools/prof.sh <tag> [bench args]   -> gpurun_out/<tag>_launches.csv, gpurun_out/<tag>.ncu-rep
tag=$1; shift
timeout -s KILL 200 ncu --metrics gpu__time_duration.sum --clock-control none -s 9 -c 12 --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 2 --warmup 3 --e2e-steps 1 --cpu-sample 16 "$@" > gpurun_out/${tag}_b.log 2>&1
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:k_inflate_ -s 6 -c 2 -o gpurun_out/${tag} python bench.py --steps 2 --warmup 3 --e2e-steps 1 --cpu-sample 16 "$@" > gpurun_out/${tag}_b2.log 2>&1
