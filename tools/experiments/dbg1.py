"""Debug driver: a few config-2 members through the C ABI, printing which path decoded them."""
import sys, time, zlib, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import threebz_b200 as t
import datagen
from tests.gpuutil import run_batch
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4
size = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
fmt = sys.argv[3] if len(sys.argv) > 3 else "zlib"
flags = int(sys.argv[4]) if len(sys.argv) > 4 else 0
ctx = t.default_ctx(0)
ms = datagen.members(n, size, 1000, fmt, threads=4)
print("launch", n, size, fmt, flags, flush=True)
t0 = time.time()
got, dev = run_batch(ctx, fmt, [c for _, c in ms], size, flags=flags)
print("done %.3f s, device %.3f ms" % (time.time() - t0, dev), flush=True)
bad = 0
for (plain, comp), g in zip(ms, got):
    ok = g["verdict"] == 0 and g["out"] == plain
    bad += not ok
paths = {}
for g in got:
    paths[g["path"]] = paths.get(g["path"], 0) + 1
print("paths", paths, "bad", bad, flush=True)
