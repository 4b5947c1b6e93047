#!/usr/bin/env python3
"""Joins an ncu SASS source page with nvdisasm line info: per source line instructions, thread
utilisation and stall samples.  usage: hotlines.py <report.ncu-rep> <lib.so> <kernel-substring> [topN]"""
import csv, io, os, re, subprocess, sys, tempfile

rep, so, kern = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, stdout=subprocess.DEVNULL)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
# walk the kernel's section: remember the current //## File "..", line N annotation per instruction
lines, cur, inside = [], None, False
for ln in dis.splitlines():
    if ln.startswith("//---") and ".text." in ln:
        inside = kern in ln
        continue
    if not inside:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", ln):
        lines.append(cur)
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "-k", "regex:" + kern], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
his = [i for i, r in enumerate(rows) if len(r) > 3 and r[0] == "Address"]
hi = his[0]
hdr = rows[hi]
ix = {h: i for i, h in enumerate(hdr)}
# (a report that holds several launches of the kernel lists each of them: the first one is used)
body = [r for r in rows[hi + 1:(his[1] if len(his) > 1 else len(rows))] if len(r) >= len(hdr) and r[0].startswith("0x")]
n = min(len(body), len(lines))
if len(body) != len(lines):
    print("warning: %d sass rows vs %d disassembled instructions" % (len(body), len(lines)), file=sys.stderr)
agg = {}
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
for r, loc in zip(body[:n], lines[:n]):
    a = agg.setdefault(loc, {"inst": 0, "thr": 0, "samp": 0, "st": {}})
    a["inst"] += int(r[ix["Instructions Executed"]]); a["thr"] += int(r[ix["Thread Instructions Executed"]])
    a["samp"] += int(r[ix["# Samples"]])
    for s in stall_cols:
        v = int(r[ix[s]] or 0)
        if v:
            a["st"][s] = a["st"].get(s, 0) + v
ti = sum(a["inst"] for a in agg.values()); ts = sum(a["samp"] for a in agg.values())
print("total warp-instructions %d, samples %d" % (ti, ts))
src_cache = {}
def src(loc):
    if not loc: return ""
    for d in ("3bz_b200/csrc",):
        p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), d, loc[0])
        if os.path.exists(p):
            if p not in src_cache: src_cache[p] = open(p).read().splitlines()
            L = src_cache[p]
            return L[loc[1] - 1].strip()[:90] if loc[1] <= len(L) else ""
    return ""
key = (lambda kv: -kv[1]["inst"]) if os.environ.get("BY") == "inst" else (lambda kv: -kv[1]["samp"])
for loc, a in sorted(agg.items(), key=key)[:top]:
    st = ",".join("%s:%d" % (k[6:], v) for k, v in sorted(a["st"].items(), key=lambda kv: -kv[1])[:3])
    print("%5.1f%% samp %5.1f%% inst thr/inst %4.1f %-22s %-40s | %s" % (100 * a["samp"] / max(1, ts), 100 * a["inst"] / max(1, ti),
          a["thr"] / max(1, a["inst"]), "%s:%d" % loc if loc else "?", st, src(loc)))
