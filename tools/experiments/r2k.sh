#!/bin/bash
# decode loop timing per member (debug build with device printf)
TBZ_LIB=$PWD/3bz_b200/var_timing.so timeout -s KILL 300 python bench.py --steps 1 --warmup 1 --e2e-steps 1 --cpu-sample 16 > gpurun_out/r2k_timing.txt 2> gpurun_out/r2k_timing.err
grep -c "^T " gpurun_out/r2k_timing.txt
