#!/bin/bash
# on the GPU box: stage breakdown of the split decode (config 3) and an ncu capture of the gzip CRC kernel (config 4)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2i_smi.txt
TBZ_KTIME=1 timeout -s KILL 300 python bench.py --workload gzip1g --steps 3 --warmup 2 --e2e-steps 1 --cpu-sample 1 --no-also > gpurun_out/r2i_gzip1g.json 2> gpurun_out/r2i_gzip1g.err
timeout -s KILL 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2i_gzip1g_launches.csv python bench.py --workload gzip1g --steps 1 --warmup 1 --e2e-steps 1 --cpu-sample 1 --no-also > /dev/null 2>&1
timeout -s KILL 400 ncu --set full --clock-control none --import-source on -k regex:k_member_crc -s 2 -c 1 -o gpurun_out/r2i_crc python bench.py --workload gzip1m --members 2048 --steps 2 --warmup 2 --e2e-steps 1 --cpu-sample 1 --no-also > gpurun_out/r2i_crc_b.log 2>&1
grep -E "tbz split" gpurun_out/r2i_gzip1g.err | tail -24
