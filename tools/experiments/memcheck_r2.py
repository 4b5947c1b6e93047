"""compute-sanitizer --tool memcheck target for the code of round 2's third session: the split decode of one member
(block-start search with per-lane tables and queues, 512-thread symbolic resolve, translate, k_span_crc / k_split_adler),
the CRC kernel on mixed lengths, and the pipelined batch path (result records stored by a kernel)."""
import zlib
import sys
sys.path.insert(0, ".")
import datagen
import threebz_b200 as t
from tests.gpuutil import run_batch

ctx = t.default_ctx(0)
plain = datagen.text(14 << 20, 5)
for fmt, ck in (("gzip", zlib.crc32(plain)), ("zlib", zlib.adler32(plain))):
    comp = datagen.compress(plain, fmt)
    got, _ = run_batch(ctx, fmt, [comp], len(plain))
    assert got[0]["path"] == 2 and got[0]["out"] == plain and got[0]["checksum"] == ck, fmt
lens = [0, 1, 17, 63, 64, 65, 1025, 65535, 65537, 200001, (1 << 20) + 1]
ps = [plain[:n] for n in lens]
got, _ = run_batch(ctx, "gzip", [datagen.compress(p, "gzip") for p in ps], lens)
assert all(g["out"] == p and g["checksum"] == zlib.crc32(p) for g, p in zip(got, ps))
ms = datagen.members(2304, 16000, 7000, "zlib")
got, _ = run_batch(ctx, "zlib", [c for _, c in ms], 16000)
assert all(g["out"] == p for g, (p, _) in zip(got, ms))
print("memcheck target ok")
