#!/bin/bash
# on the GPU box: chunk-size sweep of the split decode of one large member
for kb in ${KBS:-128 96 64 48}; do echo "chunk $kb KiB"; TBZ_SPLIT_CHUNK_KB=$kb python bench.py --workload gzip1g --steps 3 --warmup 2 --e2e-steps 1 --cpu-sample 1 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('  %.1f GB/s  %.2f ms'%(d['value'], d['ms_per_step']))"; done
