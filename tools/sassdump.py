#!/usr/bin/env python3
"""SASS of one kernel in address order with the source line of every instruction: tools/sassdump.py <lib.so> <kernel-substring>"""
import os, re, subprocess, sys, tempfile
so, kern = sys.argv[1:3]
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, stdout=subprocess.DEVNULL)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
cur, inside = None, False
for ln in dis.splitlines():
    if ln.startswith("//---") and ".text." in ln:
        inside = kern in ln; continue
    if not inside: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m: cur = "%s:%s" % (os.path.basename(m.group(1))[8:], m.group(2)); continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", ln) or re.match(r"^\.L_", ln):
        print("%-18s %s" % (cur if not ln.startswith(".L_") else "", re.sub(r"\s+", " ", ln.strip())[:110]))
