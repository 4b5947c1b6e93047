#!/bin/bash
# on the GPU box: e2e pipeline with the result records stored by a kernel (default) against the DMA (TBZ_PIPE_RESULTS_DMA=1); streams; timeline
mkdir -p gpurun_out
run() { "$@" python bench.py --no-also --steps 5 --warmup 3 --cpu-sample 16 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('  e2e %.2f GB/s %.3f ms (ceiling %.1f)  device %.1f GB/s  ok %s'%(d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e']['ceiling_gbs'], d['value'], d['verification']['ok']))"; }
for cfg in "" "TBZ_PIPE_RESULTS_DMA=1" "TBZ_PIPE_STREAMS=5" "TBZ_PIPE_STREAMS=6 TBZ_PIPE_PARTS=12" "TBZ_PIPE_STREAMS=2" ""; do
  echo "== ${cfg:-default}"
  run env $cfg
done 2>&1 | tee gpurun_out/r2pipe.log
TBZ_PIPE_TRACE=1 python bench.py --no-also --steps 3 --warmup 3 --e2e-steps 4 --cpu-sample 16 2>&1 | grep "tbz pipe" | tail -14 | tee -a gpurun_out/r2pipe.log
