#!/bin/bash
# on the GPU box: the round's parity run, bench lines, launch lists and one --set full capture -> gpurun_out/
tag=${1:-r2f}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${tag}_smi.txt 2>&1
timeout -s KILL 900 python -m pytest tests -m gpu -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${tag}_pytest.log
tail -3 gpurun_out/${tag}_pytest.log
python bench.py > gpurun_out/${tag}_bench_default.json 2> gpurun_out/${tag}_bench_default.log; echo "bench exit $?"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_bench_reference.json 2> gpurun_out/${tag}_bench_reference.log
timeout -s KILL 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 9 -c 12 --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --no-also --steps 2 --warmup 3 --e2e-steps 1 --cpu-sample 16 > gpurun_out/${tag}_b.log 2>&1
timeout -s KILL 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 12 -c 8 --csv --log-file gpurun_out/${tag}_launches_gzip1m.csv python bench.py --workload gzip1m --members 1024 --no-also --steps 2 --warmup 3 --e2e-steps 1 --cpu-sample 4 > gpurun_out/${tag}_b1m.log 2>&1
timeout -s KILL 400 ncu --set full --clock-control none --import-source on -k regex:k_inflate_ -s 6 -c 2 -o gpurun_out/${tag} python bench.py --no-also --steps 2 --warmup 3 --e2e-steps 1 --cpu-sample 16 > gpurun_out/${tag}_b2.log 2>&1
timeout -s KILL 300 compute-sanitizer --tool racecheck --racecheck-report analysis python -m pytest tests/test_gpu_parity.py -m gpu -q -k "config2_subset or edge_mix" > gpurun_out/${tag}_racecheck.txt 2>&1
tail -c 600 gpurun_out/${tag}_bench_default.json; tail -5 gpurun_out/${tag}_racecheck.txt
