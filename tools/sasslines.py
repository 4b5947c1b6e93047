#!/usr/bin/env python3
"""Static SASS instruction count per source line of one kernel (no GPU needed): tools/sasslines.py <lib.so> <kernel-substring> [file-substring]
Straight-line regions execute once per step, so this is the instruction budget of a step before any profile exists."""
import os, re, subprocess, sys, tempfile, collections
so, kern = sys.argv[1:3]
fsel = sys.argv[3] if len(sys.argv) > 3 else ""
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, stdout=subprocess.DEVNULL)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
cnt = collections.Counter(); ops = collections.defaultdict(collections.Counter)
cur, inside, total = None, False, 0
for ln in dis.splitlines():
    if ln.startswith("//---") and ".text." in ln:
        inside = kern in ln; continue
    if not inside: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m: cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
    if m:
        total += 1; cnt[cur] += 1; ops[cur][m.group(2).split(".")[0]] += 1
print("total static instructions", total)
src = {}
for (f, l), c in sorted(cnt.items(), key=lambda kv: (kv[0] or ("", 0))):
    if fsel and (not f or fsel not in f): continue
    if f not in src:
        pth = os.path.join(os.path.dirname(os.path.abspath(so)), "csrc", f)
        src[f] = open(pth).read().splitlines() if os.path.exists(pth) else []
    text = src[f][l - 1].strip()[:90] if src[f] and l <= len(src[f]) else ""
    print("%-22s %4d  %3d  %-40s | %s" % (f, l, c, " ".join("%s%d" % kv for kv in ops[(f, l)].most_common(5)), text))
