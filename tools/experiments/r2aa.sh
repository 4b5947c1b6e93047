#!/bin/bash
mkdir -p gpurun_out
# kernel times of a gzip 1 MiB batch (CRC kernel separately)
timeout -s KILL 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 12 -c 8 --csv --log-file gpurun_out/r2aa_gzip1m_launches.csv python bench.py --workload gzip1m --members 1024 --no-also --steps 2 --warmup 3 --e2e-steps 1 --cpu-sample 4 > gpurun_out/r2aa_b.log 2>&1
grep -E "k_member_crc|k_inflate" gpurun_out/r2aa_gzip1m_launches.csv | awk -F'","' '{print $5, $NF}' | head -12
# e2e: pipeline geometry
for parts in 4 8 12 24; do for streams in 3 6; do
  echo "parts $parts streams $streams: $(TBZ_PIPE_PARTS=$parts TBZ_PIPE_STREAMS=$streams timeout -s KILL 120 python bench.py --no-also --steps 3 --warmup 3 --e2e-steps 5 --cpu-sample 16 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['e2e']['value'],1), round(d['e2e']['ms_per_step'],2), 'ceiling', round(d['e2e']['ceiling_gbs'],1))")"
done; done
