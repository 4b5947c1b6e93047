#!/usr/bin/env python3
"""Aggregates profiles/hotlines.py output by the '// ----' section comments of a source file.
usage: sections.py <report.ncu-rep> <lib.so> <kernel> <source.cuh> [top]"""
import collections, os, re, subprocess, sys
rep, so, kern, srcf = sys.argv[1:5]
top = int(sys.argv[5]) if len(sys.argv) > 5 else 14
here = os.path.dirname(os.path.abspath(__file__))
out = subprocess.run([sys.executable, os.path.join(here, "..", "profiles", "hotlines.py"), rep, so, kern, "1000"],
                     capture_output=True, text=True, env=dict(os.environ, BY="inst")).stdout
rows = []
for l in out.splitlines():
    m = re.match(r'\s*([\d.]+)% samp\s+([\d.]+)% inst thr/inst\s+([\d.]+) (\S+):(\d+)', l)
    if m:
        rows.append((m.group(4), int(m.group(5)), float(m.group(2)), float(m.group(1)), l))
    elif l.startswith("total"):
        print(l)
src = open(srcf).read().splitlines()
base = os.path.basename(srcf)
marks = [(i + 1, l.strip()) for i, l in enumerate(src) if l.strip().startswith('// ----') or l.strip().startswith('// ====') or re.match(r'^(__device__|template|__global__)', l)]
def sec(f, l):
    if base not in f:
        return 'other: ' + f
    cur = 'top'
    for ln, t in marks:
        if l >= ln:
            cur = t[:60]
    return cur
agg = collections.OrderedDict()
for f, l, i, s, _ in sorted(rows, key=lambda r: (r[0], r[1])):
    a = agg.setdefault(sec(f, l), [0, 0]); a[0] += i; a[1] += s
for k, (i, s) in agg.items():
    if i > 0.2 or s > 0.2:
        print('%-62s inst %5.1f%%  samp %5.1f%%' % (k, i, s))
print("-- top by samples")
for r in sorted(rows, key=lambda r: -r[3])[:top]:
    print(r[4].rstrip()[:175])
