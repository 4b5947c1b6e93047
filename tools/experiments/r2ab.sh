#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests -m gpu -q > gpurun_out/r2ab_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2ab_pytest.log; tail -3 gpurun_out/r2ab_pytest.log
TBZ_KTIME=1 timeout -s KILL 300 python bench.py --no-also > gpurun_out/r2ab_bench.json 2> gpurun_out/r2ab_bench.err; grep "^\[tbz\]" gpurun_out/r2ab_bench.err | sed -n '8p'
timeout -s KILL 600 compute-sanitizer --tool racecheck --racecheck-report analysis python -m pytest tests/test_gpu_parity.py -m gpu -q -k "config2_subset or config4_members or edge_mix" > gpurun_out/r2ab_racecheck.txt 2>&1
grep -E "Race reported|hazards\]|RACECHECK SUMMARY|passed|failed" gpurun_out/r2ab_racecheck.txt | sed 's/=========//' | cut -c1-190 | sort | uniq -c | sort -rn | head -20
timeout -s KILL 600 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py -m gpu -q -k "config2_subset or config4_members or edge_mix or truncation or corruption" > gpurun_out/r2ab_memcheck.txt 2>&1
tail -4 gpurun_out/r2ab_memcheck.txt
