#!/bin/bash
# parity + the default bench line (headline + also) of the working tree
tag=${1:-r2t}
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${tag}_pytest.log
tail -5 gpurun_out/${tag}_pytest.log
SECONDS=0; timeout -s KILL 900 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench exit $?"
tail -c 3000 gpurun_out/${tag}_bench.json; echo "bench wall ${SECONDS}s"
