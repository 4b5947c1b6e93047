"""GPU parity: the CUDA path through the C ABI vs the CPU oracle, bit-exact (integer/byte work)."""
import hashlib
import random
import zlib

import pytest

import datagen
from tests import cases
from tests.gpuutil import compare, run_batch

pytestmark = pytest.mark.gpu


def test_config1_fixture(engine, ctx, oracle):
    """3bz:decompress-vector on test.deflated (:format :deflate :start 8) — bit-exact output + size."""
    raw, meta = cases.test_deflated()
    out = bytearray(meta["size_field"])
    buf, n = engine.decompress_vector(raw, format="deflate", start=8, output=out)
    assert n == 22728 and hashlib.sha256(bytes(buf[:n])).hexdigest() == meta["sha256"]
    buf, n = engine.decompress_vector(raw, format=":deflate", start=8)       # no :output (api.lisp:50-65)
    assert n == 22728 and hashlib.sha256(bytes(buf)).hexdigest() == meta["sha256"]
    with pytest.raises(engine.ThreeBzError, match="not enough space"):
        engine.decompress_vector(raw, format="deflate", start=8, output=bytearray(22727))
    with pytest.raises(engine.ThreeBzError, match="incomplete"):
        engine.decompress_vector(raw, format="deflate", start=8, end=len(raw) - 1, output=bytearray(22728))


def test_nayuki_vectors(ctx, oracle):
    vs = cases.nayuki()
    ins = [bytes.fromhex(v["input_hex"]) for v in vs]
    got, _ = run_batch(ctx, "deflate", ins, 1024)
    for v, d, g in zip(vs, ins, got):
        w = oracle.decompress_vector(d, "deflate", out_cap=1024)
        compare(g, w, v["line"])
        if v["marker"] is None:
            assert g["verdict"] == 0 and g["out"].hex() == v["expected_hex"]


def test_edge_mix(ctx, oracle):
    for fmt in cases.FMTS:
        items = [(n, c, p) for n, f, c, p in cases.edge_streams() if f == fmt]
        got, _ = run_batch(ctx, fmt, [c for _, c, _ in items], [len(p) for _, _, p in items])
        for (name, comp, plain), g in zip(items, got):
            w = oracle.decompress_vector(comp, fmt, out_cap=len(plain))
            compare(g, w, (name, fmt))
            assert g["verdict"] == 0 and g["out"] == plain, (name, fmt)


def test_truncation_sweep(ctx, oracle):
    """every prefix of a stream: same verdict, same bytes so far (deflate.lisp:114-120, :399-427)."""
    raw, meta = cases.test_deflated()
    payload = raw[8:]
    ins = [payload[:k] for k in range(0, len(payload) + 1, 1)]
    got, _ = run_batch(ctx, "deflate", ins, 22728)
    for k, g in enumerate(got):
        compare(g, oracle.decompress_vector(ins[k], "deflate", out_cap=22728), k)
    for fmt in ("zlib", "gzip"):
        plain = datagen.text(3000, 11) + bytes(500) + datagen.random_bytes(300, 3)
        comp = datagen.compress(plain, fmt)
        ins = [comp[:k] for k in range(len(comp) + 1)]
        got, _ = run_batch(ctx, fmt, ins, len(plain))
        for k, g in enumerate(got):
            w = oracle.decompress_vector(ins[k], fmt, out_cap=len(plain))
            compare(g, w, (fmt, k))
            if w["verdict"] == 1:
                st = oracle.State(fmt, output_size=len(plain))
                ret = st.decompress(st.make_context(ins[k]))
                want_where = 1 if ret == w["out_len"] and w["out_len"] else None
                if want_where:
                    assert g["where"] in (1, 2) or fmt == "zlib"


def test_capacity_sweep(ctx, oracle):
    """output-overflow fires exactly when the buffer is full (deflate.lisp:239-241,254-269,693-697)."""
    for name, fmt, comp, plain in cases.edge_streams():
        if name not in ("text64k", "zeros", "stored", "fixed", "period3") or fmt == "gzip":
            continue
        n = len(plain)
        caps = sorted(set([0, 1, 2, 3, 100, 257, 258, 259, 4095, 4096, n // 2, n - 259, n - 2, n - 1, n, n + 1, n + 100]))
        caps = [c for c in caps if c >= 0]
        got, _ = run_batch(ctx, fmt, [comp] * len(caps), caps)
        for c, g in zip(caps, got):
            compare(g, oracle.decompress_vector(comp, fmt, out_cap=c), (name, fmt, c))


def test_corruption_fuzz(ctx, oracle):
    """random bit flips / byte stomps: verdict class and produced bytes equal the oracle's."""
    rnd = random.Random(2024)
    bases = []
    for fmt in cases.FMTS:
        bases.append((fmt, datagen.compress(datagen.text(6000, 21), fmt), 6000))
        bases.append((fmt, datagen.compress(datagen.text(4000, 22), fmt, strategy=zlib.Z_FIXED), 4000))
        bases.append((fmt, datagen.compress(datagen.random_bytes(3000, 23), fmt, level=0), 3000))
    for fmt, comp, n in bases:
        ins = []
        for _ in range(400):
            b = bytearray(comp)
            for _ in range(rnd.choice((1, 1, 2, 5))):
                k = rnd.randrange(len(b)) if rnd.random() < .7 else rnd.randrange(min(len(b), 80))
                if rnd.random() < .5:
                    b[k] ^= 1 << rnd.randrange(8)
                else:
                    b[k] = rnd.randrange(256)
            ins.append(bytes(b))
        cap = n + 5000
        got, _ = run_batch(ctx, fmt, ins, cap)
        for i, g in enumerate(got):
            compare(g, oracle.decompress_vector(ins[i], fmt, out_cap=cap), (fmt, i))


def test_odd_huffman_codes(ctx, oracle):
    """hand-built dynamic headers: lone-symbol codes of every length, HLIT > 286, unused distance tree."""
    def bits_to_bytes(bits):
        out = bytearray((len(bits) + 7) // 8)
        for i, c in enumerate(bits):
            if c == "1":
                out[i // 8] |= 1 << (i % 8)
        return bytes(out)

    ins = []
    rnd = random.Random(7)
    for _ in range(300):
        # random (mostly invalid) dynamic headers followed by random payload bits
        b = "1" + "01" + "".join(rnd.choice("01") for _ in range(rnd.randrange(14, 600)))
        ins.append(bits_to_bytes(b))
    got, _ = run_batch(ctx, "deflate", ins, 4096)
    for i, g in enumerate(got):
        compare(g, oracle.decompress_vector(ins[i], "deflate", out_cap=4096), i)


def test_config2_subset(ctx, oracle):
    """batch of independent 64 KiB zlib members (level 6, dynamic Huffman) — bit-exact, Adler verdicts."""
    ms = datagen.members(96, 65536, 1000, "zlib")
    got, _ = run_batch(ctx, "zlib", [c for _, c in ms], 65536)
    for i, ((plain, comp), g) in enumerate(zip(ms, got)):
        assert g["verdict"] == 0 and g["out"] == plain and g["checksum"] == zlib.adler32(plain), i
        assert g["in_used"] == len(comp)
        assert g["path"] == 1, "clean members must be decoded by the fast lane-parallel kernel"
    w = oracle.decompress_vector(ms[0][1], "zlib", out_cap=65536)
    compare(got[0], w, 0)
    # the same members through the sequential kernel only
    got2, _ = run_batch(ctx, "zlib", [c for _, c in ms], 65536, flags=2)
    for (plain, comp), g in zip(ms, got2):
        assert g["verdict"] == 0 and g["out"] == plain and g["path"] == 0


def test_config4_members(ctx, oracle):
    """1 MiB gzip members, several dynamic blocks each."""
    ms = datagen.members(6, 1 << 20, 5000, "gzip")
    got, _ = run_batch(ctx, "gzip", [c for _, c in ms], 1 << 20)
    for (plain, comp), g in zip(ms, got):
        assert g["verdict"] == 0 and g["out"] == plain and g["checksum"] == zlib.crc32(plain)
        assert g["path"] == 1


@pytest.mark.gpu
def test_gzip_crc_span_geometry(ctx, oracle):
    """k_member_crc (inflate_crc.cuh): member lengths around every boundary of its segment geometry (1 024 threads, 64-byte
    runs, the last segment on thread 0), packed back to back so that the device outputs start at every alignment, mixed
    lengths in one launch (the cached weights are recomputed between members), and CRCs that disagree with the trailer."""
    lens = [0, 1, 2, 15, 16, 17, 63, 64, 65, 127, 1023, 1024, 1025, 4095, 65535, 65536, 65537, 65536 + 64, 65599,
            70001, 2 * 65536 - 1, 200000, 65536, 65536, 1, 1 << 20, (1 << 20) + 1, (1 << 20) - 1, 3 << 20, 1234567]
    text = datagen.text(max(lens), 4242)
    plain = [text[:n] for n in lens]
    comp = [datagen.compress(p, "gzip") for p in plain]
    for flags in (0,):
        got, _ = run_batch(ctx, "gzip", comp, [len(p) for p in plain], flags=flags)
        for p, g in zip(plain, got):
            assert g["verdict"] == 0 and g["out"] == p and g["checksum"] == zlib.crc32(p) and g["path"] == 1, len(p)
    bad = [bytearray(c) for c in comp[5:]]
    for b in bad:
        b[-6] ^= 0x40
    got, _ = run_batch(ctx, "gzip", [bytes(b) for b in bad], [len(p) for p in plain[5:]])
    for b, p, g in zip(bad, plain[5:], got):
        compare(g, oracle.decompress_vector(bytes(b), "gzip", out_cap=len(p)), ("bad crc", len(p)))


def test_gzip_header_fields(ctx, oracle):
    plain = datagen.text(5000, 5)
    variants = [cases.gzip_with_header_fields(plain),
                cases.gzip_with_header_fields(plain, extra=None),
                cases.gzip_with_header_fields(plain, name=None, comment=None, hcrc=False),
                cases.gzip_with_header_fields(plain, extra=b"", name=b"", comment=b"")]
    bad = bytearray(variants[0]); bad[12] ^= 0x40
    variants.append(bytes(bad))
    variants += [variants[0][:k] for k in range(0, 60)]
    got, _ = run_batch(ctx, "gzip", variants, 5000)
    for i, g in enumerate(got):
        compare(g, oracle.decompress_vector(variants[i], "gzip", out_cap=5000), i)
    # name / comment / extra are skipped by the fast kernels (path 1); a header CRC goes to the sequential one
    assert [got[i]["path"] for i in range(4)] == [0, 0, 1, 0], [got[i]["path"] for i in range(4)]
    fast = [cases.gzip_with_header_fields(plain, hcrc=False), cases.gzip_with_header_fields(plain, extra=None, hcrc=False),
            cases.gzip_with_header_fields(plain, extra=b"", name=b"", comment=b"", hcrc=False)]
    got, _ = run_batch(ctx, "gzip", fast, 5000)
    for i, g in enumerate(got):
        compare(g, oracle.decompress_vector(fast[i], "gzip", out_cap=5000), ("fast", i))
        assert g["verdict"] == 0 and g["out"] == plain and g["path"] == 1, ("fast", i, g["path"])


def _drain_both(engine, oracle, comp, fmt, sizes, start=0):
    sizes = list(sizes)
    it = iter(sizes)
    ost = oracle.State(fmt, output_size=sizes[0])
    octx = ost.make_context(comp, start=start)
    first = bytearray(next(it))
    est = {"deflate": engine.make_deflate_state, "zlib": engine.make_zlib_state,
           "gzip": engine.make_gzip_state}[fmt](output_buffer=first)
    ectx = engine.make_octet_vector_context(comp, start=start)
    out = bytearray()
    for _ in range(len(sizes)):
        want = ost.decompress(octx)
        got = engine.decompress(ectx, est)
        assert got == want
        assert (engine.finished(est), engine.input_underrun(est), engine.output_overflow(est)) == \
               (ost.finished, ost.input_underrun, ost.output_overflow)
        assert bytes(est.output_buffer[:got]) == ost.output(want)
        if ost.finished or ost.output_overflow:
            out += est.output_buffer[:got]
            if ost.finished:
                return bytes(out)
            nxt = next(it)
            ost.replace_output_buffer(nxt)
            engine.replace_output_buffer(est, bytearray(nxt))
        else:
            return bytes(out)
    raise AssertionError("ran out of buffers")


def test_chunked_output_sessions(engine, ctx, oracle):
    """config 5: drained through replace-output-buffer; per-call return values and flags equal."""
    raw, meta = cases.test_deflated()
    ref = zlib.decompress(raw[8:], -15)
    assert _drain_both(engine, oracle, raw, "deflate", [3] * (len(ref) // 3 + 2), start=8) == ref
    rnd = random.Random(77)
    for _ in range(5):
        sizes = [1 + rnd.randrange(12345) for _ in range(len(ref) + 2)]
        assert _drain_both(engine, oracle, raw, "deflate", sizes, start=8) == ref
    for name, fmt, comp, plain in cases.edge_streams():
        if name not in ("stored", "zeros", "rle", "period3", "fixed", "empty", "one_byte", "text64k"):
            continue
        n = len(plain)
        for sizes in ([32768] * (n // 32768 + 2), [max(1, n)] * 2,
                      [1 + rnd.randrange(12345) for _ in range(n // 1 + 2)][: n + 2]):
            assert _drain_both(engine, oracle, comp, fmt, sizes) == plain, (name, fmt)


def test_chunked_input_sessions(engine, ctx, oracle):
    """§8f.1: input in pieces; after each call finished or input-underrun (test-chunked-input.lisp:27-44)."""
    raw, meta = cases.test_deflated()
    payload = raw[8:]
    ref = zlib.decompress(payload, -15)
    rnd = random.Random(5)
    for gen in (lambda: 997, lambda: 1 + rnd.randrange(1233)):
        ost = oracle.State("deflate", output_size=len(ref))
        est = engine.make_deflate_state(output_buffer=bytearray(len(ref)))
        o = 0
        while o < len(payload):
            end = min(len(payload), o + gen())
            want = ost.decompress(ost.make_context(payload[o:end]))
            got = engine.decompress(engine.make_octet_vector_context(payload[o:end]), est)
            assert got == want
            assert (engine.finished(est), engine.input_underrun(est)) == (ost.finished, ost.input_underrun)
            assert bytes(est.output_buffer[:got]) == ost.output(want)
            o = end
        assert engine.finished(est) and bytes(est.output_buffer[:got]) == ref


def test_chunked_input_three_octets_at_a_time(engine, ctx, oracle):
    """test-chunked-input.lisp:27-44: the fixture fed in 3-octet pieces; every call's return value and flags are the
    oracle's.  (The session keeps what it was given on the device and uploads only the new octets.)"""
    import time
    raw, meta = cases.test_deflated()
    payload = raw[8:]
    t0 = time.time()
    ost = oracle.State("deflate", output_size=22728)
    est = engine.make_deflate_state(output_buffer=bytearray(22728))
    calls = 0
    for o in range(0, len(payload), 3):
        piece = payload[o:o + 3]
        want = ost.decompress(ost.make_context(piece))
        got = engine.decompress(engine.make_octet_vector_context(piece), est)
        calls += 1
        assert got == want, (o, got, want)
        assert (engine.finished(est), engine.input_underrun(est), engine.output_overflow(est)) == \
               (ost.finished, ost.input_underrun, ost.output_overflow), o
        if ost.finished:
            break
    assert ost.finished and bytes(est.output_buffer[:got]) == ost.output(want)
    assert time.time() - t0 < 10.0, "chunked input must stay cheap (%d calls took %.1f s)" % (calls, time.time() - t0)


@pytest.mark.parametrize("fmt", ["gzip", "zlib", "deflate"])
def test_chunked_input_resumes_at_block_boundaries(engine, ctx, oracle, fmt):
    """a stream of many blocks fed in 16 KiB pieces: every call's return value and flags are the oracle's, and a call
    costs the block it is in, not the stream so far (the session resumes at the last block boundary it reached;
    deflate.lisp:65-88,114-137 saves its whole state instead)"""
    import time
    plain = datagen.text(6 << 20, 21)
    comp = datagen.compress(plain, fmt)
    ost = oracle.State(fmt, output_size=len(plain) + 10)
    est = {"deflate": engine.make_deflate_state, "zlib": engine.make_zlib_state,
           "gzip": engine.make_gzip_state}[fmt](output_buffer=bytearray(len(plain) + 10))
    t0 = time.time()
    got = want = 0
    for o in range(0, len(comp), 16384):
        piece = comp[o:o + 16384]
        want = ost.decompress(ost.make_context(piece))
        got = engine.decompress(engine.make_octet_vector_context(piece), est)
        assert got == want, (o, got, want)
        assert (engine.finished(est), engine.input_underrun(est), engine.output_overflow(est)) == \
               (ost.finished, ost.input_underrun, ost.output_overflow), o
    assert ost.finished and got == len(plain) and bytes(est.output_buffer[:got]) == plain
    assert time.time() - t0 < 30.0, "%.1f s" % (time.time() - t0)


def test_context_stops_behind_the_stream(engine, ctx, oracle):
    """the context's offset after a stream finished is just past its last octet (io.lisp:17-58): trailing data or the
    next member starts there"""
    p1, c1 = datagen.member(30000, 11, "gzip")
    p2, c2 = datagen.member(20000, 12, "gzip")
    data = c1 + c2 + b"tail"
    ectx = engine.make_octet_vector_context(data)
    est = engine.make_gzip_state(output_buffer=bytearray(40000))
    n = engine.decompress(ectx, est)
    assert engine.finished(est) and bytes(est.output_buffer[:n]) == p1
    assert ectx.offset == len(c1), (ectx.offset, len(c1))
    est2 = engine.make_gzip_state(output_buffer=bytearray(40000))
    n2 = engine.decompress(ectx, est2)
    assert engine.finished(est2) and bytes(est2.output_buffer[:n2]) == p2 and ectx.offset == len(c1) + len(c2)


def test_decompress_batch_api(engine, ctx):
    ms = datagen.members(8, 20000, 300, "zlib")
    ins = [c for _, c in ms]
    ins[3] = ins[3][:-3]            # a bad member never poisons the batch
    res = engine.decompress_batch(ins, "zlib", 20000)
    for i, (buf, n, v) in enumerate(res):
        assert bytes(buf[:n]) == ms[i][0]
        assert v == (1 if i == 3 else 0)


def test_octet_pointer_context(engine, ctx):
    import ctypes as C
    plain, comp = datagen.member(50000, 8, "gzip")
    raw = C.create_string_buffer(comp, len(comp))
    with engine.with_octet_pointer(C.addressof(raw), len(comp)) as op:
        st = engine.make_gzip_state(output_buffer=bytearray(50000))
        n = engine.decompress(engine.make_octet_pointer_context(op), st)
        assert engine.finished(st) and bytes(st.output_buffer[:n]) == plain
    with pytest.raises(engine.ThreeBzError):
        engine.decompress(engine.make_octet_pointer_context(op), engine.make_gzip_state(output_buffer=bytearray(8)))


def test_split_single_member(engine, ctx, oracle):
    """config 3 shape (scaled down): ONE large member is cut at fixed offsets, block starts are found
    speculatively, chunks are decoded in parallel and unresolved window references patched afterwards."""
    plain = datagen.text(24 << 20, 3)
    for fmt in ("gzip", "zlib", "deflate"):
        comp = datagen.compress(plain, fmt)
        got, _ = run_batch(ctx, fmt, [comp], len(plain))
        g = got[0]
        assert g["path"] == 2, "large members must take the split path"
        assert g["verdict"] == 0 and g["out_len"] == len(plain)
        assert hashlib.sha256(g["out"]).digest() == hashlib.sha256(plain).digest()
        ck = {"gzip": zlib.crc32(plain), "zlib": zlib.adler32(plain), "deflate": 0}[fmt]
        assert g["checksum"] == ck and g["in_used"] == len(comp)
        # the same stream with the split disabled: identical result record
        got2, _ = run_batch(ctx, fmt, [comp], len(plain), flags=4)
        assert got2[0]["path"] != 2 and got2[0]["out"] == g["out"] and got2[0]["checksum"] == g["checksum"]
    # a gzip header with optional fields (`gzip file` always writes a name; gzip.lisp:178-260): still the split path
    for kw in ({"extra": None, "comment": None, "hcrc": False}, {}, {"extra": b"x" * 5000, "hcrc": False}):
        comp = cases.gzip_with_header_fields(plain, **kw)
        got, _ = run_batch(ctx, "gzip", [comp], len(plain))
        g = got[0]
        assert g["path"] == 2, ("a gzip member with header fields must take the split path", kw)
        assert g["verdict"] == 0 and g["out_len"] == len(plain) and g["checksum"] == zlib.crc32(plain) and g["in_used"] == len(comp)
        assert hashlib.sha256(g["out"]).digest() == hashlib.sha256(plain).digest()
    # damage in the middle, a wrong trailer, a truncated stream, too small a buffer: verdict and bytes as the oracle's
    comp = datagen.compress(plain[: 6 << 20], "gzip")
    n = 6 << 20
    bad_mid = bytearray(comp); bad_mid[len(comp) // 2] ^= 0x55
    bad_crc = bytearray(comp); bad_crc[-6] ^= 1
    variants = [bytes(bad_mid), bytes(bad_crc), comp[: len(comp) - 5], comp[: len(comp) // 3]]
    got, _ = run_batch(ctx, "gzip", variants, n)
    for v, g in zip(variants, got):
        compare(g, oracle.decompress_vector(v, "gzip", out_cap=n), "split variant")
    got, _ = run_batch(ctx, "gzip", [comp], n - 1000)
    compare(got[0], oracle.decompress_vector(comp, "gzip", out_cap=n - 1000), "split overflow")


def test_gzip_concatenated_members(engine, ctx, oracle):
    """New entry point: the members of a multi-member gzip file, back to back; the default
    one-member behaviour (gzip.lisp:279-286) is what decompress_vector keeps doing."""
    parts = [datagen.text(40000, 71), b"", datagen.text(3, 72), datagen.text(150000, 73)]
    blobs = [datagen.compress(p, "gzip") for p in parts]
    data = b"".join(blobs)
    out = bytearray(sum(map(len, parts)))
    res, used = engine.decompress_gzip_members(data, out, ctx=ctx)
    assert used == len(data) and len(res) == len(parts)
    assert bytes(out) == b"".join(parts)
    for (n, iu, ck, v), p, b in zip(res, parts, blobs):
        assert (n, iu, ck, v) == (len(p), len(b), zlib.crc32(p), 0)
        w = oracle.decompress_vector(b, "gzip", out_cap=max(1, len(p)))
        assert w["verdict"] == 0 and w["checksum"] == ck and w["out_len"] == n
    # a damaged third member stops the walk there and reports its verdict; the first two stand
    bad = bytearray(data)
    bad[len(blobs[0]) + len(blobs[1]) + len(blobs[2]) - 8] ^= 1          # CRC-32 of member 2
    res, used = engine.decompress_gzip_members(bytes(bad), out, ctx=ctx)
    assert len(res) == 3 and [r[3] for r in res[:2]] == [0, 0] and res[2][3] == 32 and used == len(blobs[0]) + len(blobs[1])
    # the reference-compatible call still stops after the first member
    buf, n = engine.decompress_vector(data, format="gzip", output=bytearray(len(out)))
    assert n == len(parts[0]) and bytes(buf[:n]) == parts[0]


@pytest.mark.parametrize("fmt", ["deflate", "zlib", "gzip"])
def test_device_pointers_any_alignment(ctx, oracle, fmt):
    """Caller-owned device memory at every input / output misalignment: the fast kernels' byte-wise
    flush (output not 16-byte aligned) and word-misaligned bit readers, through TBZ_FLAG_DEVICE_PTRS."""
    import ctypes as C
    from threebz_b200 import _ffi
    L = _ffi.lib()
    size = 70000
    plains = [datagen.text(size, 4200 + i) for i in range(16)]
    comps = [datagen.compress(p, fmt) for p in plains]
    stride_in = max(map(len, comps)) + 64
    stride_out = size + 64
    d_in, d_out = C.c_void_p(), C.c_void_p()
    _ffi.check(L.tbz_device_alloc(ctx.h, 16 * stride_in, C.byref(d_in)), ctx.h)
    _ffi.check(L.tbz_device_alloc(ctx.h, 16 * stride_out, C.byref(d_out)), ctx.h)
    try:
        marr = (_ffi.Member * 16)()
        for i, c in enumerate(comps):
            a_in = d_in.value + i * stride_in + i            # input misalignment 0..15
            a_out = d_out.value + i * stride_out + (i * 7) % 16   # output misalignment, all residues mod 16
            _ffi.check(L.tbz_memcpy_h2d(ctx.h, a_in, c, len(c)), ctx.h)
            marr[i] = _ffi.Member(a_in, len(c), a_out, size)
        rarr = (_ffi.Result * 16)()
        _ffi.check(L.tbz_inflate_batch(ctx.h, _ffi.fmt_code(fmt), marr, 16, rarr, _ffi.FLAG_DEVICE_PTRS, None), ctx.h)
        for i, (p, c) in enumerate(zip(plains, comps)):
            w = oracle.decompress_vector(c, fmt, out_cap=size)
            host = C.create_string_buffer(size)
            _ffi.check(L.tbz_memcpy_d2h(ctx.h, host, marr[i].out, size), ctx.h)
            r = rarr[i]
            assert (r.verdict, r.out_len, r.checksum, r.in_used) == (0, size, w["checksum"], len(c)), (i, r.verdict, r.out_len)
            assert host.raw == p, i
            assert r.path == 1, (i, r.path)                  # the fast kernels, not the sequential one
    finally:
        L.tbz_device_free(ctx.h, d_in)
        L.tbz_device_free(ctx.h, d_out)


def test_stored_blocks_take_the_fast_path(ctx, oracle):
    """Stored blocks (deflate.lisp:532-573) travel through the token kernels as literal tokens:
    level-0 streams, incompressible data (libz falls back to stored blocks), streams that mix stored
    and Huffman blocks (Z_FULL_FLUSH emits an empty stored block), in all three formats."""
    text, rnd = datagen.text(200000, 5), datagen.random_bytes(150000, 9)
    for fmt in cases.FMTS:
        mixed = zlib.compressobj(6, zlib.DEFLATED, datagen.WBITS[fmt])
        blob = mixed.compress(text[:70000]) + mixed.flush(zlib.Z_FULL_FLUSH) + mixed.compress(rnd[:80000]) + \
            mixed.flush(zlib.Z_SYNC_FLUSH) + mixed.compress(text[70000:]) + mixed.flush()
        plain_mixed = text[:70000] + rnd[:80000] + text[70000:]
        items = [(datagen.compress(text, fmt, level=0), text), (datagen.compress(rnd, fmt), rnd), (blob, plain_mixed),
                 (datagen.compress(b"", fmt, level=0), b""), (datagen.compress(b"x", fmt, level=0), b"x")]
        got, _ = run_batch(ctx, fmt, [c for c, _ in items], [max(1, len(p)) for _, p in items])
        for k, ((comp, plain), g) in enumerate(zip(items, got)):
            compare(g, oracle.decompress_vector(comp, fmt, out_cap=max(1, len(plain))), (fmt, k))
            assert g["verdict"] == 0 and g["out"] == plain and g["path"] == 1, (fmt, k, g["verdict"], g["path"])
        # a stored block cut short, a bad NLEN and a too small buffer keep the reference's verdicts (sequential kernel)
        comp = datagen.compress(rnd[:5000], fmt, level=0)
        hdr = {"deflate": 0, "zlib": 2, "gzip": 10}[fmt]
        bad = bytearray(comp); bad[hdr + 3] ^= 0x40
        ins = [comp[:hdr + 1000], bytes(bad), comp]
        got, _ = run_batch(ctx, fmt, ins, [5000, 5000, 4999])
        for k, g in enumerate(got):
            compare(g, oracle.decompress_vector(ins[k], fmt, out_cap=[5000, 5000, 4999][k]), (fmt, "bad", k))
        assert [g["verdict"] for g in got] == [1, 17, 2], [g["verdict"] for g in got]


def test_octet_stream_context(engine, ctx, tmp_path):
    """octet-stream-context (io-common.lisp:47-63, io.lisp:61-104): a member that lies
    somewhere inside a file, read through a stream context; %resync-file-stream puts the file
    position where the context says."""
    import io
    plain = datagen.text(300000, 77)
    comp = datagen.compress(plain, "gzip")
    blob = b"HEAD" * 25 + comp + b"TAIL" * 10
    path = tmp_path / "blob.bin"
    path.write_bytes(blob)
    for f in (open(path, "rb"), io.BytesIO(blob)):
        with f:
            st = engine.make_gzip_state(output_buffer=bytearray(len(plain)))
            c = engine.make_octet_stream_context(f, start=100, offset=100, end=100 + len(comp))
            n = engine.decompress(c, st)
            assert n == len(plain) and engine.finished(st) and bytes(st.output_buffer[:n]) == plain
            engine.resync_file_stream(c)
            assert f.tell() == c.offset == 100 + len(comp) and f.read(4) == b"TAIL"
            # chunked input through the stream: first half, then the rest (test-chunked-input.lisp:27-44)
            st = engine.make_gzip_state(output_buffer=bytearray(len(plain)))
            half = 100 + len(comp) // 2
            engine.decompress(engine.make_octet_stream_context(f, start=100, offset=100, end=half), st)
            assert engine.input_underrun(st) and not engine.finished(st)
            n = engine.decompress(engine.make_octet_stream_context(f, start=half, offset=half, end=100 + len(comp)), st)
            assert n == len(plain) and engine.finished(st) and bytes(st.output_buffer[:n]) == plain
    with pytest.raises(engine.ThreeBzError):
        engine.make_octet_stream_context(object())


def test_output_tail_is_untouched(engine, ctx):
    """the bytes of an output buffer past the returned count are the caller's (api.lisp:35-61 never writes them):
    oversized, pre-filled buffers — one member (every decompress-vector :output call), adjacent members (the direct
    DMA path), a batch large enough for the pipelined path, and one large enough for its small leading parts (>= 2 048
    members)"""
    import ctypes as C
    from threebz_b200 import _ffi
    L = _ffi.lib()
    for n, size in ((1, 5000), (7, 3000), (600, 60000), (2304, 16000)):
        ms = datagen.members(n, size, 7000, "zlib")
        cap = size + 1000 + 17
        blob = b"".join(c for _, c in ms)
        inbuf = (C.c_uint8 * len(blob)).from_buffer_copy(blob)
        outbuf = (C.c_uint8 * (cap * n))()
        C.memset(outbuf, 0xA5, cap * n)
        marr = (_ffi.Member * n)()
        io = 0
        for i, (p, c) in enumerate(ms):
            marr[i] = _ffi.Member(C.addressof(inbuf) + io, len(c), C.addressof(outbuf) + i * cap, cap)
            io += len(c)
        rarr = (_ffi.Result * n)()
        _ffi.check(L.tbz_inflate_batch(ctx.h, _ffi.fmt_code("zlib"), marr, n, rarr, 0, None), ctx.h)
        raw = bytes(outbuf)
        for i, (p, c) in enumerate(ms):
            assert rarr[i].verdict == 0 and rarr[i].out_len == size
            assert raw[i * cap:i * cap + size] == p, (n, i)
            assert raw[i * cap + size:(i + 1) * cap] == b"\xa5" * (cap - size), ("tail overwritten", n, i)


def test_batch_multi_partitions_over_contexts(engine, ctx, oracle):
    """tbz_inflate_batch_multi, the product's host-side partitioner: members of mixed sizes and formats'
    verdicts are spread over several engine contexts — one per GPU where the box has several, two contexts on the one
    GPU otherwise (the partitioning, the per-device threads and the gathering of results are the same code) — and every
    result equals the oracle's, in the caller's order."""
    ng = engine.device_count()
    ctxs = [engine.Ctx(d) for d in range(min(ng, 4))] if ng > 1 else [engine.Ctx(0), engine.Ctx(0)]
    rnd = random.Random(3)
    ms = [datagen.member(rnd.choice((3000, 20000, 65536, 200000)), 900 + i, "zlib") for i in range(70)]
    comps = [c for _, c in ms]
    comps[5] = comps[5][: len(comps[5]) // 2]                      # truncated
    bad = bytearray(comps[9]); bad[len(bad) // 2] ^= 0x10; comps[9] = bytes(bad)
    caps = [len(p) for p, _ in ms]
    caps[11] -= 100                                               # too small an output
    res = engine.decompress_batch(comps, "zlib", caps, ctxs=ctxs)
    for i, (buf, count, verdict) in enumerate(res):
        want = oracle.decompress_vector(comps[i], "zlib", out_cap=caps[i])
        assert verdict == want["verdict"] and count == want["out_len"], (i, verdict, want["verdict"])
        assert bytes(buf[:count]) == want["out"], i
    # members of similar size: consecutive ranges (the direct-DMA, pipelined path per device)
    ms = datagen.members(1200, 30000, 4000, "zlib")
    res = engine.decompress_batch([c for _, c in ms], "zlib", 30000, ctxs=ctxs)
    for i, (buf, count, verdict) in enumerate(res):
        assert verdict == 0 and count == 30000 and bytes(buf) == ms[i][0], i
    for c in ctxs:
        c.close()
