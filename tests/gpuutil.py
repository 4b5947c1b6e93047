"""Helpers for the -m gpu parity tests: everything goes through the C ABI (ctypes)."""
import ctypes as C

import threebz_b200 as t
from threebz_b200 import _ffi


def run_batch(ctx, fmt, inputs, caps, flags=0):
    """inputs: list of bytes; caps: int or list.  Returns list of dicts like the oracle's."""
    L = _ffi.lib()
    n = len(inputs)
    caps = [caps] * n if isinstance(caps, int) else list(caps)
    blob = b"".join(inputs)
    inbuf = (C.c_uint8 * max(1, len(blob))).from_buffer_copy(blob if blob else b"\0")
    outbuf = (C.c_uint8 * max(1, sum(caps)))()
    marr = (_ffi.Member * max(1, n))()
    io = oo = 0
    offs = []
    for i, (d, c) in enumerate(zip(inputs, caps)):
        marr[i] = _ffi.Member(C.addressof(inbuf) + io, len(d), C.addressof(outbuf) + oo, c)
        offs.append(oo)
        io += len(d)
        oo += c
    rarr = (_ffi.Result * max(1, n))()
    ms = C.c_float()
    _ffi.check(L.tbz_inflate_batch(ctx.h, _ffi.fmt_code(fmt), marr, n, rarr, flags, C.byref(ms)), ctx.h)
    mv = memoryview(outbuf)
    res = []
    for i in range(n):
        r = rarr[i]
        res.append({"verdict": r.verdict, "out_len": r.out_len, "checksum": r.checksum, "where": r.where,
                    "in_used": r.in_used, "path": r.path,
                    "out": bytes(mv[offs[i]:offs[i] + r.out_len])})
    return res, ms.value


def compare(got, want, what=""):
    assert got["verdict"] == want["verdict"], (what, got["verdict"], want["verdict"], got["out_len"], want["out_len"])
    assert got["out_len"] == want["out_len"], (what, got["out_len"], want["out_len"])
    assert got["out"] == want["out"], what
    assert got["checksum"] == want["checksum"], (what, hex(got["checksum"]), hex(want["checksum"]))
