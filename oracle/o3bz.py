"""ctypes binding of the parity oracle (oracle/oracle3bz.c).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module; the product package (3bz_b200) never does.
"""
import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle3bz.so")

DEFLATE, ZLIB, GZIP = 0, 1, 2
FORMATS = {"deflate": DEFLATE, "zlib": ZLIB, "gzip": GZIP}
FINISHED, INPUT_UNDERRUN, OUTPUT_OVERFLOW = 0, 1, 2
ERR_BUFFER_SWITCH = 33


def build(force=False):
    src = os.path.join(_HERE, "oracle3bz.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "liboracle3bz.so"])
    return _SO


class Stats(C.Structure):
    _fields_ = [("total_out", C.c_uint64), ("literals", C.c_uint64), ("matches", C.c_uint64),
                ("match_bytes", C.c_uint64), ("stored_bytes", C.c_uint64),
                ("blocks", C.c_uint64 * 3), ("header_bits", C.c_uint64)]

    def as_dict(self):
        return {"total_out": self.total_out, "literals": self.literals, "matches": self.matches,
                "match_bytes": self.match_bytes, "stored_bytes": self.stored_bytes,
                "blocks": list(self.blocks), "header_bits": self.header_bits}


class Context(C.Structure):
    _fields_ = [("p", C.c_void_p), ("start", C.c_size_t), ("end", C.c_size_t), ("offset", C.c_size_t)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        L.o3bz_state_new.restype = C.c_void_p
        L.o3bz_state_new.argtypes = [C.c_int]
        L.o3bz_state_free.argtypes = [C.c_void_p]
        L.o3bz_set_output.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
        L.o3bz_replace_output_buffer.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
        L.o3bz_context_init.argtypes = [C.POINTER(Context), C.c_void_p, C.c_size_t, C.c_size_t]
        L.o3bz_decompress.restype = C.c_int64
        L.o3bz_decompress.argtypes = [C.POINTER(Context), C.c_void_p]
        for f in ("o3bz_finished", "o3bz_input_underrun", "o3bz_output_overflow", "o3bz_error"):
            getattr(L, f).argtypes = [C.c_void_p]
        L.o3bz_checksum.restype = C.c_uint32
        L.o3bz_checksum.argtypes = [C.c_void_p]
        L.o3bz_get_stats.argtypes = [C.c_void_p, C.POINTER(Stats)]
        L.o3bz_decompress_vector.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, C.c_int, C.c_void_p,
                                             C.c_size_t, C.POINTER(C.c_size_t), C.POINTER(C.c_uint32),
                                             C.POINTER(Stats)]
        L.o3bz_decompress_vector_grow.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, C.c_int,
                                                  C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]
        L.o3bz_free.argtypes = [C.c_void_p]
        L.o3bz_adler32.argtypes = [C.c_void_p, C.c_size_t, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
        L.o3bz_crc32.restype = C.c_uint32
        L.o3bz_crc32.argtypes = [C.c_void_p, C.c_size_t, C.c_uint32]
        L.o3bz_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                 C.c_void_p, C.c_int, C.c_size_t, C.c_size_t]
        C.cast(L.o3bz_state_new(0), C.c_void_p)  # builds the static tables once, before any thread
        _lib = L
    return _lib


def _fmt(f):
    return FORMATS[f] if isinstance(f, str) else int(f)


def _buf(b):
    """bytes-like -> (ctypes array keeping it alive, address)."""
    if isinstance(b, (bytes, bytearray, memoryview)):
        arr = (C.c_uint8 * max(1, len(b))).from_buffer_copy(bytes(b) if len(b) else b"\0")
        return arr, C.addressof(arr)
    raise TypeError(type(b))


def decompress_vector(data, fmt="zlib", start=0, end=None, out_cap=None):
    """api.lisp:36-48 with :output of out_cap bytes.  Returns dict(verdict, out, checksum, stats)."""
    L = lib()
    end = len(data) if end is None else end
    keep, addr = _buf(data)
    if out_cap is None:
        out_cap = 1 << 16
        while True:  # find a capacity that does not overflow (test convenience, not the Lisp path)
            r = decompress_vector(data, fmt, start, end, out_cap)
            if r["verdict"] != OUTPUT_OVERFLOW:
                return r
            out_cap *= 4
    out = (C.c_uint8 * max(1, out_cap))()
    n = C.c_size_t(0)
    ck = C.c_uint32(0)
    st = Stats()
    v = L.o3bz_decompress_vector(addr, start, end, _fmt(fmt), out, out_cap, C.byref(n), C.byref(ck), C.byref(st))
    return {"verdict": v, "out": bytes(out[:n.value]) if n.value else b"", "out_len": n.value,
            "checksum": ck.value, "stats": st.as_dict()}


def decompress_vector_grow(data, fmt="zlib", start=0, end=None):
    """api.lisp:50-65 (no :output): returns (verdict, bytes)."""
    L = lib()
    end = len(data) if end is None else end
    keep, addr = _buf(data)
    p = C.c_void_p()
    n = C.c_size_t(0)
    v = L.o3bz_decompress_vector_grow(addr, start, end, _fmt(fmt), C.byref(p), C.byref(n))
    out = C.string_at(p, n.value) if p and n.value else b""
    if p:
        L.o3bz_free(p)
    return v, out


class State:
    """make-deflate-state / make-zlib-state / make-gzip-state (+ :output-buffer)."""

    def __init__(self, fmt, output_size=None):
        self.L = lib()
        self.h = C.c_void_p(self.L.o3bz_state_new(_fmt(fmt)))
        self.buf = None
        if output_size is not None:
            self.buf = (C.c_uint8 * max(1, output_size))()
            self.cap = output_size
            self.L.o3bz_set_output(self.h, self.buf, output_size)
        self._ctx_keep = []

    def __del__(self):
        if getattr(self, "h", None):
            self.L.o3bz_state_free(self.h)
            self.h = None

    def replace_output_buffer(self, size):
        buf = (C.c_uint8 * max(1, size))()
        r = self.L.o3bz_replace_output_buffer(self.h, buf, size)
        if r:
            raise RuntimeError("can't switch buffers without filling old one yet.")
        self.buf, self.cap = buf, size

    def make_context(self, data, start=0, end=None):
        keep, addr = _buf(data)
        c = Context()
        self.L.o3bz_context_init(C.byref(c), addr, start, len(data) if end is None else end)
        self._ctx_keep.append(keep)
        return c

    def decompress(self, ctx):
        """Returns the Lisp return value, or raises RuntimeError(code) where the Lisp signals."""
        r = self.L.o3bz_decompress(C.byref(ctx), self.h)
        if r < 0:
            raise RuntimeError(self.L.o3bz_error(self.h))
        return r

    finished = property(lambda s: bool(s.L.o3bz_finished(s.h)))
    input_underrun = property(lambda s: bool(s.L.o3bz_input_underrun(s.h)))
    output_overflow = property(lambda s: bool(s.L.o3bz_output_overflow(s.h)))
    error = property(lambda s: s.L.o3bz_error(s.h))
    checksum = property(lambda s: s.L.o3bz_checksum(s.h))

    def output(self, n):
        return bytes(self.buf[:n])

    def stats(self):
        st = Stats()
        self.L.o3bz_get_stats(self.h, C.byref(st))
        return st.as_dict()


def adler32(data, s1=1, s2=0):
    keep, addr = _buf(data)
    a, b = C.c_uint32(s1), C.c_uint32(s2)
    lib().o3bz_adler32(addr, len(data), C.byref(a), C.byref(b))
    return a.value | (b.value << 16)


def crc32(data, crc=0):
    keep, addr = _buf(data)
    return lib().o3bz_crc32(addr, len(data), crc)


def batch(inputs, fmt, caps, threads=1):
    """Decodes many members inside C (no per-member Python overhead), `threads` host threads.
    Returns (seconds, verdicts, out_lens, match_bytes, outputs-arena, offsets)."""
    import time
    from concurrent.futures import ThreadPoolExecutor
    L = lib()
    n = len(inputs)
    caps = [caps] * n if isinstance(caps, int) else list(caps)
    blob = b"".join(inputs)
    inbuf = (C.c_uint8 * max(1, len(blob))).from_buffer_copy(blob if blob else b"\0")
    outbuf = (C.c_uint8 * max(1, sum(caps)))()
    ip, il = (C.c_void_p * n)(), (C.c_size_t * n)()
    op, oc = (C.c_void_p * n)(), (C.c_size_t * n)()
    ol, vd, mb = (C.c_size_t * n)(), (C.c_int * n)(), (C.c_uint64 * n)()
    io = oo = 0
    offs = []
    for i in range(n):
        ip[i], il[i], op[i], oc[i] = C.addressof(inbuf) + io, len(inputs[i]), C.addressof(outbuf) + oo, caps[i]
        offs.append(oo)
        io += len(inputs[i]); oo += caps[i]
    threads = max(1, min(threads, n))
    per = (n + threads - 1) // threads
    t = time.perf_counter()
    with ThreadPoolExecutor(threads) as ex:
        list(ex.map(lambda k: L.o3bz_batch(ip, il, op, oc, ol, vd, mb, _fmt(fmt), k * per, min(n, (k + 1) * per)),
                    range(threads)))
    dt = time.perf_counter() - t
    return dt, list(vd), list(ol), list(mb), outbuf, offs
