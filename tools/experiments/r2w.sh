#!/bin/bash
# parity + headline bench + ncu (--set full) of both hot kernels
tag=${1:-r2w}
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${tag}_pytest.log
tail -4 gpurun_out/${tag}_pytest.log
TBZ_KTIME=1 timeout -s KILL 300 python bench.py --no-also > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; grep "^\[tbz\]" gpurun_out/${tag}_bench.err | sed -n '8p'; python -c "
import json;d=json.load(open('gpurun_out/${tag}_bench.json'));print(d['value'],d['ms_per_step'],d['e2e']['value'],d['sustained']['value'])"
timeout -s KILL 400 ncu --set full --clock-control none --import-source on -k regex:k_inflate_ -s 6 -c 2 -o gpurun_out/${tag} python bench.py --no-also --steps 2 --warmup 3 --e2e-steps 1 --cpu-sample 16 > gpurun_out/${tag}_ncu.log 2>&1
ls -la gpurun_out | grep ${tag}
