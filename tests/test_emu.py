"""Kernel logic on the CPU: the batched kernels' device source (3bz_b200/csrc/*.cuh, the same text nvcc compiles
for sm_100a) runs under tests/emu — a SIMT emulator with fibers for threads, checked warp collectives and
-fsanitize=alignment — and is compared with the oracle.  The -m gpu suite repeats all of this on the B200
through the C ABI; this suite exists so that a kernel bug is found before GPU time is spent.
Test infrastructure only: the product has no CPU path."""
import random
import zlib

import pytest

import datagen
from oracle import o3bz
from tests import cases, emuutil
from tests.gpuutil import compare

VARIANTS = [0]          # (one pair of batched kernels: huff_decode.cuh + lz_resolve.cuh)


@pytest.fixture(scope="module", autouse=True)
def _built():
    emuutil.build()


@pytest.mark.parametrize("variant", VARIANTS)
def test_emu_config2_members(variant):
    ms = datagen.members(6, 65536, 1000, "zlib")
    got, nseq = emuutil.run_batch("zlib", [c for _, c in ms], 65536, variant=variant)
    assert nseq == 0
    for (p, c), g in zip(ms, got):
        compare(g, o3bz.decompress_vector(c, "zlib", out_cap=65536), "cfg2")
        assert g["out"] == p and g["path"] == 1 and g["checksum"] == zlib.adler32(p)


def test_emu_config4_member():
    p, c = datagen.member(1 << 20, 5000, "gzip")
    got, nseq = emuutil.run_batch("gzip", [c], [len(p)])
    assert nseq == 0 and got[0]["out"] == p and got[0]["checksum"] == zlib.crc32(p) and got[0]["path"] == 1


@pytest.mark.parametrize("fmt", cases.FMTS)
def test_emu_edge_mix(fmt):
    items = [(n, c, p) for n, f, c, p in cases.edge_streams() if f == fmt]
    got, nseq = emuutil.run_batch(fmt, [c for _, c, _ in items], [len(p) for _, _, p in items])
    for (name, comp, plain), g in zip(items, got):
        compare(g, o3bz.decompress_vector(comp, fmt, out_cap=len(plain)), (name, fmt))
        assert g["verdict"] == 0 and g["out"] == plain and g["path"] == 1, (name, fmt, g["path"])


def test_emu_nayuki_vectors():
    vs = cases.nayuki()
    ins = [bytes.fromhex(v["input_hex"]) for v in vs]
    got, _ = emuutil.run_batch("deflate", ins, 1024)
    for v, d, g in zip(vs, ins, got):
        compare(g, o3bz.decompress_vector(d, "deflate", out_cap=1024), v["line"])


def test_emu_fixture_and_truncations():
    raw, meta = cases.test_deflated()
    payload = raw[8:]
    ks = sorted(set(list(range(0, 40)) + list(range(40, len(payload), 97)) + list(range(len(payload) - 12, len(payload) + 1))))
    ins = [payload[:k] for k in ks]
    got, _ = emuutil.run_batch("deflate", ins, 22728)
    for k, d, g in zip(ks, ins, got):
        compare(g, o3bz.decompress_vector(d, "deflate", out_cap=22728), k)
    assert got[-1]["out_len"] == 22728 and got[-1]["path"] == 1


def test_emu_capacity_sweep():
    """an output buffer that is too small: the fast kernels give the member up, the sequential kernel reports
    output-overflow with exactly the bytes that fit (deflate.lisp:239-241, :254-269, :693-697)"""
    for name, fmt, comp, plain in cases.edge_streams():
        if name not in ("text64k", "zeros", "period3") or fmt != "zlib":
            continue
        n = len(plain)
        caps = sorted(set([0, 1, 3, 257, 258, 259, 4096, n // 2, n - 259, n - 1, n, n + 1, n + 100]))
        got, _ = emuutil.run_batch(fmt, [comp] * len(caps), caps)
        for cap, g in zip(caps, got):
            compare(g, o3bz.decompress_vector(comp, fmt, out_cap=cap), (name, cap))


def test_emu_any_alignment():
    """members and outputs at every 16-byte residue (caller-owned device memory need not be aligned)"""
    p, c = datagen.member(40000, 321, "zlib")
    z = datagen.compress(bytes(5000) + p[:3000] + b"ab" * 3000, "zlib")
    zp = bytes(5000) + p[:3000] + b"ab" * 3000
    for mis in (1, 2, 3, 4, 7, 8, 13, 15):
        got, nseq = emuutil.run_batch("zlib", [c, z], [len(p), len(zp)], in_mis=mis, out_mis=(mis * 5) % 16)
        assert nseq == 0
        assert got[0]["out"] == p and got[1]["out"] == zp and got[0]["path"] == 1 and got[1]["path"] == 1


def test_emu_corruption_fuzz():
    """mutated streams: whatever the fast kernels make of them, the verdict and bytes are the oracle's"""
    rnd = random.Random(5)
    p, c = datagen.member(20000, 99, "zlib")
    ins = []
    for _ in range(60):
        b = bytearray(c)
        for _ in range(rnd.randint(1, 3)):
            b[rnd.randrange(len(b))] ^= 1 << rnd.randrange(8)
        ins.append(bytes(b))
    got, _ = emuutil.run_batch("zlib", ins, 20000)
    for d, g in zip(ins, got):
        compare(g, o3bz.decompress_vector(d, "zlib", out_cap=20000), "fuzz")


def test_emu_long_distances_and_lengths():
    """distances up to 32 KiB (older than the 16 KiB ring: read back from the output) and 258-byte matches"""
    rnd = random.Random(7)
    block = bytes(rnd.randrange(256) for _ in range(33000))
    plain = block + block[:32768][::-1][:100] + block[100:30000] + block[:258] * 3 + block[5:32700]
    for level in (6, 9):
        comp = datagen.compress(plain, "zlib", level=level)
        got, nseq = emuutil.run_batch("zlib", [comp], [len(plain)])
        assert nseq == 0 and got[0]["out"] == plain and got[0]["path"] == 1


@pytest.mark.parametrize("out_mis", [0, 1, 7, 15])
def test_emu_gzip_crc_span_geometry(out_mis):
    """k_member_crc (inflate_crc.cuh): lengths around every boundary of its segment geometry (NT = 1024 threads, 64-byte
    runs, the 16-byte alignment offset, the last segment on thread 0), at several output misalignments; mixed lengths in
    one launch, so the cached weights are recomputed between members."""
    lens = [0, 1, 2, 15, 16, 17, 63, 64, 65, 127, 1023, 1024, 1025, 4095, 65535, 65536, 65537, 65536 + 64, 65600 - out_mis,
            70001, 2 * 65536 - 1, 200000, 65536, 65536, 1]
    text = datagen.text(max(lens), 4242)
    plain = [text[:n] for n in lens]
    comp = [datagen.compress(p, "gzip") for p in plain]
    got, nseq = emuutil.run_batch("gzip", comp, [len(p) for p in plain], out_mis=out_mis)
    assert nseq == 0
    for p, c, g in zip(plain, comp, got):
        assert g["verdict"] == 0 and g["out"] == p and g["checksum"] == zlib.crc32(p) and g["path"] == 1, len(p)
    # a damaged CRC goes to the sequential kernel, which reports it like the oracle
    bad = [bytearray(c) for c in comp[5:12]]
    for b in bad:
        b[-6] ^= 0x40
    got, nseq = emuutil.run_batch("gzip", [bytes(b) for b in bad], [len(p) for p in plain[5:12]], out_mis=out_mis)
    assert nseq == len(bad)
    for b, p, g in zip(bad, plain[5:12], got):
        compare(g, o3bz.decompress_vector(bytes(b), "gzip", out_cap=len(p)), ("bad crc", len(p)))


def test_emu_block_start_search():
    """The split decode's block-start search (inflate_decode.cuh find_block_start: two filters, survivors validated one
    per lane).  Ground truth without a bit-level parser: a Z_FULL_FLUSH ends the stream on a byte boundary behind an
    empty stored block, so the next block starts at a known bit; on text it is a dynamic block."""
    text = datagen.text(600000, 31337)
    co = zlib.compressobj(6, zlib.DEFLATED, -15)
    pieces, starts, pos = [], [], 0
    for k in range(6):
        starts.append(pos * 8)                       # a block starts here (the first one, or the one behind a flush)
        d = co.compress(text[k * 100000:(k + 1) * 100000]) + (co.flush(zlib.Z_FULL_FLUSH) if k < 5 else co.flush())
        pieces.append(d)
        pos += len(d)
    raw = b"".join(pieces)
    assert zlib.decompress(raw, -15) == text
    for k in range(1, 5):
        s = starts[k]
        assert (raw[s // 8] & 7) == 4, "not a non-final dynamic block: the test's assumption about libz is off"
        # searched from the start itself, from inside the flush marker in front of it, and from far in front of it: libz
        # cuts 100 000 bytes of text into several blocks, so the hit from far away is some earlier real start
        assert emuutil.find_block_start(raw, s, s + 4096) == s
        assert emuutil.find_block_start(raw, s - 30, s + 4096) == s
        far = emuutil.find_block_start(raw, starts[k - 1] + 1, s + 64)
        assert far is not None and starts[k - 1] < far <= s
        # nothing between a start and the end of its header
        assert emuutil.find_block_start(raw, s + 1, s + 200) is None
    # the last piece: its first block starts behind the flush; the stream's final block is never reported (BFINAL = 1)
    s = starts[5]
    first = emuutil.find_block_start(raw, s, len(raw) * 8)
    assert first == s or (raw[s // 8] & 1) == 1
    # random bits: no hit
    rnd = datagen.random_bytes(1 << 16, 99)
    assert emuutil.find_block_start(rnd, 0, len(rnd) * 8 - 64) is None


def test_emu_split_adler():
    """k_split_adler (zlib members of the split decode): 64 KiB segments, 16-byte loads, every alignment class"""
    data = datagen.random_bytes(300001, 5) + bytes(70000) + b"\xff" * 70000
    for n in (0, 1, 15, 16, 17, 65535, 65536, 65537, 131072, 200000, len(data)):
        for mis in (0, 1, 8, 15):
            assert emuutil.split_adler(data[:n], mis) == zlib.adler32(data[:n]), (n, mis)


def test_emu_tail_compose_and_translate():
    """k_tail_compose (markers of map v looked up in map v - stride) and k_split_translate (symbols to bytes; a marker is a
    byte of the 32 KiB in front of its chunk) against numpy"""
    import numpy as np
    rng = np.random.default_rng(7)
    MARK = 0x8000
    nv = 5
    maps = rng.integers(0, 256, size=(nv, 32768), dtype=np.uint16)
    mark = rng.random((nv, 32768)) < 0.3
    maps[mark] = MARK | rng.integers(0, 32768, size=int(mark.sum()), dtype=np.uint16)
    for stride in (1, 2, 4):
        want = maps.copy()
        for v in range(stride, nv):
            m = (maps[v] & MARK) != 0
            want[v][m] = maps[v - stride][maps[v][m] & 0x7fff]
        assert np.array_equal(emuutil.tail_compose(maps, stride), want), stride
    # translate: three chunks behind 40 000 bytes of known output; the tails (last 32 KiB of a chunk) are final already
    sizes = [40000, 50001, 33000, 70017]
    offs = np.cumsum([0] + sizes)
    total = int(offs[-1])
    final = rng.integers(0, 256, size=total, dtype=np.uint8)
    # markers of chunk k sit in front of its own tail and point into the 32 KiB in front of the chunk — the tail of
    # chunk k - 1, which is final (literal bytes) before the translation starts
    out = bytearray(total)
    sym = final.astype(np.uint16)
    final2 = final.copy()
    for k in range(1, len(sizes)):
        a, b = int(offs[k]), int(offs[k + 1])
        body = np.arange(a, max(a, b - 32768))           # the part of the chunk in front of its own tail
        if body.size:
            pos = rng.choice(body, size=body.size // 3, replace=False)
            idx = rng.integers(0, 32768, size=pos.size)
            sym[pos] = MARK | idx
            final2[pos] = final2[a - 32768 + idx]        # (the window in front of chunk k is a final tail: literal bytes)
    for k in range(len(sizes)):
        a, b = int(offs[k]), int(offs[k + 1])
        t0 = max(a, b - 32768)
        out[t0:b] = final2[t0:b].tobytes()               # k_tail_write's work
    for mis in (0, 3):
        got = emuutil.split_translate(offs, sym, out, grid=3, mis=mis)
        assert got == final2.tobytes(), mis
