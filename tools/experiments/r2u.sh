#!/bin/bash
# on the GPU box: chunk-size sweep of the split decode, 1 GiB and 256 MiB members
mkdir -p gpurun_out
for wl in gzip1g gzip256m; do
for kb in ${KBS:-512 384 320 256 224 160}; do echo "$wl chunk $kb KiB"; TBZ_SPLIT_CHUNK_KB=$kb python bench.py --workload $wl --steps 3 --warmup 2 --e2e-steps 1 --cpu-sample 1 --no-also 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('  %.1f GB/s  %.2f ms'%(d['value'], d['ms_per_step']))"; done
done 2>&1 | tee gpurun_out/r2u_split.log
TBZ_SPLIT_CHUNK_KB=320 TBZ_KTIME=1 timeout -s KILL 300 python bench.py --workload gzip1g --steps 2 --warmup 2 --e2e-steps 1 --cpu-sample 1 --no-also 2>&1 | grep "tbz split" | tail -7 | tee -a gpurun_out/r2u_split.log
