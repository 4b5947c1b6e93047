#!/bin/bash
# on the GPU box: occupancy variants of phase two (threads per CTA), then the round's final evidence run
mkdir -p gpurun_out
bash tools/sweep.sh --no-also > gpurun_out/r2z_sweep.log 2>&1
grep -E "^==|^\[tbz\]" gpurun_out/r2z_sweep.log
rm -f 3bz_b200/var_*.so
bash tools/final_r2.sh r2z
