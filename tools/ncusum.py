#!/usr/bin/env python3
"""Summarises gpurun_out/<tag>_launches.csv and <tag>.ncu-rep: per-kernel time, key metrics."""
import csv, subprocess, sys, io
tag = sys.argv[1]
try:
    rows = list(csv.reader(l for l in open("gpurun_out/%s_launches.csv" % tag) if l.startswith('"')))
    h = rows[0]; ki = h.index("Kernel Name"); vi = h.index("Metric Value")
    agg = {}
    for r in rows[1:]:
        agg.setdefault(r[ki].split("(")[0], []).append(float(r[vi].replace(",", "")))
    for k, v in agg.items():
        print("%-22s n=%d avg %.3f ms" % (k, len(v), sum(v) / len(v) / 1e6))
except Exception as e:
    print("no launch list:", e)
out = subprocess.run(["ncu", "-i", "gpurun_out/%s.ncu-rep" % tag, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[0]
want = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'launch__registers_per_thread', 'launch__shared_mem_per_block_dynamic', 'sm__inst_executed.avg.per_cycle_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio', 'launch__grid_size',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum']
for w in want:
    if w in hdr:
        i = hdr.index(w)
        print(w, rows[1][i], '|', ' | '.join(r[i][:34] for r in rows[2:]))
