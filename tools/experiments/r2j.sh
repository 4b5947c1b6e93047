#!/bin/bash
# on the GPU box: parity tests of the default build, then per-kernel times of every variant (tools/sweep.sh)
mkdir -p gpurun_out
tag=${1:-r2j}
timeout -s KILL 900 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; tail -3 gpurun_out/${tag}_pytest.log
bash tools/sweep.sh --no-also > gpurun_out/${tag}_sweep.log 2>&1
grep -E "^==|^\[tbz\]" gpurun_out/${tag}_sweep.log
