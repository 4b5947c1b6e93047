#!/bin/bash
# on the GPU box: the round's bench lines, launch list and one --set full capture -> gpurun_out/
tag=$1
python bench.py > gpurun_out/${tag}_bench_zlib64k.json 2> gpurun_out/${tag}_bench_zlib64k.log
python bench.py --workload gzip1m --steps 5 --warmup 3 > gpurun_out/${tag}_bench_gzip1m.json 2> gpurun_out/${tag}_bench_gzip1m.log
python bench.py --workload gzip1g --steps 5 --warmup 3 > gpurun_out/${tag}_bench_gzip1g.json 2> gpurun_out/${tag}_bench_gzip1g.log
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_bench_reference.json 2> gpurun_out/${tag}_bench_reference.log
timeout -s KILL 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 9 -c 12 --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 2 --warmup 3 --e2e-steps 1 --cpu-sample 16 > gpurun_out/${tag}_b.log 2>&1
timeout -s KILL 400 ncu --set full --clock-control none --import-source on -k regex:k_inflate_ -s 6 -c 2 -o gpurun_out/${tag} python bench.py --steps 2 --warmup 3 --e2e-steps 1 --cpu-sample 16 > gpurun_out/${tag}_b2.log 2>&1
tail -c 400 gpurun_out/${tag}_bench_zlib64k.json
