"""Host-side mirror of 3bz's public API (package.lisp:13-27, api.lisp) over libthreebz_cuda.so.

The reference's host language is Common Lisp; no Lisp implementation exists in this image, so
lisp/ holds the real CFFI shim (unrunnable here) and this module issues the SAME call
sequences through ctypes, with the same names, argument meaning and error behaviour:

    decompress_vector(compressed, format=:zlib, start=0, end=None, output=None) -> (buffer, count)
    make_deflate_state / make_zlib_state / make_gzip_state (output_buffer=...)
    make_octet_vector_context(vector, start=0, offset=start, end=len)
    with_octet_pointer(pointer, size) + make_octet_pointer_context(op, start=0, offset=0, end=size)
    decompress(context, state) -> offset ; finished / input_underrun / output_overflow
    replace_output_buffer(state, buffer)
    decompress_batch(members, format, capacities) -> [(buffer, count, verdict)]   (new entry point)

Everything runs on the GPU engine; nothing here falls back to a CPU decoder.
"""
import ctypes as C
import threading

from . import _ffi
from ._ffi import EngineError, check, fmt_code, lib


class ThreeBzError(Exception):
    """Stands for the plain Lisp `error` / `assert` / `ecase` conditions the reference signals."""

    def __init__(self, message, verdict=None):
        super().__init__(message)
        self.verdict = verdict


_tls = threading.local()


def default_ctx(device=0):
    """One engine context per (thread, device), created on first use."""
    key = "ctx%d" % device
    ctx = getattr(_tls, key, None)
    if ctx is None:
        ctx = Ctx(device)
        setattr(_tls, key, ctx)
    return ctx


class Ctx:
    def __init__(self, device=0):
        self.L = lib()
        h = C.c_void_p()
        check(self.L.tbz_ctx_create(device, 0, C.byref(h)))
        self.h = h
        self.device = device

    def close(self):
        if self.h:
            self.L.tbz_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def launches(self):
        n = C.c_uint64()
        check(self.L.tbz_ctx_launch_count(self.h, C.byref(n)), self.h)
        return n.value


def _addr(buf):
    """Address of a writable/readable bytes-like without copying (bytes are copied once)."""
    if isinstance(buf, bytes):
        keep = (C.c_uint8 * max(1, len(buf))).from_buffer_copy(buf if buf else b"\0")
        return keep, C.addressof(keep)
    if isinstance(buf, (bytearray, memoryview)):
        mv = memoryview(buf)
        if mv.readonly:
            keep = (C.c_uint8 * max(1, len(mv))).from_buffer_copy(bytes(mv) if len(mv) else b"\0")
        else:
            keep = (C.c_uint8 * len(mv)).from_buffer(mv) if len(mv) else (C.c_uint8 * 1)()
        return keep, C.addressof(keep)
    if isinstance(buf, C.Array):
        return buf, C.addressof(buf)
    raise TypeError("expected an octet vector (bytes / bytearray / ctypes array), got %r" % type(buf))


# ---- contexts (io-common.lisp:36-45, io-mmap.lisp:21-54) -----------------------------------
class OctetVectorContext:
    def __init__(self, vector, start=0, offset=None, end=None):
        self.vector = vector
        self.start = start
        self.offset = start if offset is None else offset
        self.end = len(vector) if end is None else end
        self._keep, self._base = _addr(vector)

    def _unread(self):
        return self._base + self.offset, self.end - self.offset


def make_octet_vector_context(vector, start=0, offset=None, end=None):
    return OctetVectorContext(vector, start, offset, end)


class OctetPointer:
    """with-octet-pointer (io-mmap.lisp:26-40): valid only inside the `with` block.  The range is
    registered with the driver so the engine DMAs straight from it."""

    def __init__(self, pointer, size):
        self.base, self.size, self.valid, self._registered = int(pointer), int(size), False, False

    def __enter__(self):
        self.valid = True
        if self.size > 0 and lib().tbz_host_register(self.base, self.size, 0) == 0:
            self._registered = True
        return self

    def __exit__(self, *a):
        self.valid = False
        if self._registered:
            lib().tbz_host_unregister(self.base)
            self._registered = False


def with_octet_pointer(pointer, size):
    return OctetPointer(pointer, size)


class OctetPointerContext:
    def __init__(self, op, start=0, offset=0, end=None):
        self.op, self.start, self.offset = op, start, offset
        self.end = op.size if end is None else end

    def _unread(self):
        if not (self.op.valid and self.op.base and self.op.size > 0):   # io-mmap.lisp:36-40,64
            raise ThreeBzError("octet-pointer used outside its dynamic extent")
        return self.op.base + self.offset, self.end - self.offset


def make_octet_pointer_context(octet_pointer, start=0, offset=0, end=None):
    return OctetPointerContext(octet_pointer, start, offset, end)


class OctetStreamContext:
    """octet-stream-context (io-common.lisp:47-63, io.lisp:61-104): the octets [offset, end) of a
    seekable binary stream.  The reference pulls 4/8 octets at a time through FILE-POSITION and
    READ-BYTE ("very slow", README.md:13); here the unread octets are read into one host buffer per
    `decompress` call and travel like an octet vector."""

    def __init__(self, stream, start=0, offset=0, end=None):
        if not (hasattr(stream, "read") and hasattr(stream, "seek") and getattr(stream, "readable", lambda: True)()):
            raise ThreeBzError("not a valid octet stream")            # (assert (valid-octet-stream stream)), io.lisp:69
        self.stream, self.start, self.offset = stream, start, offset
        if end is None:                                               # (file-length file-stream)
            here = stream.tell()
            end = stream.seek(0, 2)
            stream.seek(here)
        self.end = end
        self._keep = None

    def _unread(self):
        self.stream.seek(self.offset)
        data = self.stream.read(max(0, self.end - self.offset))
        self._keep, base = _addr(data)
        return base, len(data)


def make_octet_stream_context(file_stream, start=0, offset=0, end=None):
    return OctetStreamContext(file_stream, start, offset, end)


def resync_file_stream(context):
    """%resync-file-stream (io-common.lisp:51-56): put the stream where the context's offset is."""
    if isinstance(context, OctetStreamContext):
        context.stream.seek(context.offset)


# ---- states (deflate.lisp:4-62, zlib.lisp:3-12, gzip.lisp:3-28) ------------------------------
class _State:
    format = None

    def __init__(self, output_buffer=None, ctx=None):
        self.ctx = ctx or default_ctx()
        self.L = lib()
        h = C.c_void_p()
        check(self.L.tbz_session_create(self.ctx.h, fmt_code(self.format), C.byref(h)), self.ctx.h)
        self.h = h
        self.output_buffer = None
        self._keep = None
        self.output_offset = 0
        if output_buffer is not None:
            self._set(output_buffer)

    def _set(self, buf):
        self._keep, addr = _addr(buf)
        check(self.L.tbz_session_set_output(self.h, addr, len(buf)), self.ctx.h)
        self.output_buffer = buf
        self.output_offset = 0

    def __del__(self):
        try:
            if self.h:
                self.L.tbz_session_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def _flags(self):
        f, u, o = C.c_int32(), C.c_int32(), C.c_int32()
        check(self.L.tbz_session_flags(self.h, C.byref(f), C.byref(u), C.byref(o)), self.ctx.h)
        return bool(f.value), bool(u.value), bool(o.value)


class DeflateState(_State):
    format = "deflate"


class ZlibState(_State):
    format = "zlib"


class GzipState(_State):
    format = "gzip"


def make_deflate_state(output_buffer=None, ctx=None):
    return DeflateState(output_buffer, ctx)


def make_zlib_state(output_buffer=None, ctx=None):
    return ZlibState(output_buffer, ctx)


def make_gzip_state(output_buffer=None, ctx=None):
    return GzipState(output_buffer, ctx)


def decompress(context, state):
    """api.lisp:3-10.  Returns the current offset into the output buffer."""
    addr, n = context._unread()
    ret, verdict = C.c_int64(), C.c_int32()
    if state.output_buffer is None:
        state._set(bytearray(0))               # (make-array 0), deflate.lisp:47
    rc = state.L.tbz_session_decompress(state.h, addr if n else None, n, C.byref(ret), C.byref(verdict))
    # the session owns every unread octet — unless the stream finished inside them: the context then stops just past
    # the consumed octets (io.lisp:17-58), where trailing data or the next member starts
    used = C.c_uint64(n)
    if rc == 0:
        state.L.tbz_session_consumed(state.h, C.byref(used))
    context.offset = min(context.end, context.offset + used.value)
    if rc == _ffi.E_STATE:
        raise ThreeBzError("decompress called on a finished or failed state", verdict.value)
    check(rc, state.ctx.h)
    if verdict.value >= 16:
        raise ThreeBzError(_ffi.verdict_name(verdict.value), verdict.value)
    state.output_offset = ret.value
    return ret.value


def replace_output_buffer(state, buffer):
    """api.lisp:12-21."""
    keep, addr = _addr(buffer)
    rc = state.L.tbz_session_replace_output(state.h, addr, len(buffer))
    if rc == _ffi.E_BUFFER_SWITCH:
        raise ThreeBzError("can't switch buffers without filling old one yet.")
    check(rc, state.ctx.h)
    state._keep, state.output_buffer, state.output_offset = keep, buffer, 0


def finished(state):
    return state._flags()[0]


def input_underrun(state):
    return state._flags()[1]


def output_overflow(state):
    return state._flags()[2]


# ---- decompress-vector (api.lisp:23-65) ---------------------------------------------------------
def _fmt_name(format):
    return format.lstrip(":") if isinstance(format, str) else {0: "deflate", 1: "zlib", 2: "gzip"}[format]


def decompress_vector(compressed, format="zlib", start=0, end=None, output=None, ctx=None):
    """Returns (buffer, count).  With OUTPUT: one decode into it, errors as the Lisp words them.
    Without: the engine sizes the result (the Lisp's doubling-buffer loop collapses to one
    device-side growth loop; the returned vector and length are the same)."""
    ctx = ctx or default_ctx()
    L = lib()
    end = len(compressed) if end is None else end
    keep, base = _addr(compressed)
    res = _ffi.Result()
    name = _fmt_name(format)
    if output is not None:
        okeep, oaddr = _addr(output)
        check(L.tbz_inflate_single(ctx.h, fmt_code(format), base + start, end - start, oaddr, len(output),
                                   C.byref(res), 0, None), ctx.h)
        if isinstance(output, bytes):
            raise TypeError("output must be writable")
        if res.verdict != _ffi.FINISHED:
            if res.verdict == _ffi.INPUT_UNDERRUN:
                raise ThreeBzError("incomplete %s stream" % name, res.verdict)
            if res.verdict == _ffi.OUTPUT_OVERFLOW:
                raise ThreeBzError("not enough space to decompress %s stream" % name, res.verdict)
            raise ThreeBzError(_ffi.verdict_name(res.verdict), res.verdict)
        return output, res.out_len
    p = C.c_void_p()
    check(L.tbz_inflate_alloc(ctx.h, fmt_code(format), base + start, end - start, C.byref(p), C.byref(res)), ctx.h)
    try:
        if res.verdict != _ffi.FINISHED:
            if res.verdict == _ffi.INPUT_UNDERRUN:       # (assert (not (ds-input-underrun state)))
                raise ThreeBzError("incomplete %s stream" % name, res.verdict)
            raise ThreeBzError(_ffi.verdict_name(res.verdict), res.verdict)
        buf = bytearray(C.string_at(p, res.out_len)) if res.out_len else bytearray()
    finally:
        L.tbz_free(p)
    return buf, len(buf)


# ---- decompress-batch: the new entry point for many independent members -------------------------
_GZ_OS = ("fat", "amiga", "vms", "unix", "vm/cms", "atari-tos", "hpfs", "macintosh", "z-system", "cp/m",
          "tops-20", "ntfs", "qdos", "acorn-riscos")          # gzip.lisp:169-176


def gzip_header(compressed, start=0, end=None):
    """The gzip-state slots decompress-gzip fills from the member header (gzip.lisp:17-28, :113-260):
    flags, extra, name, comment, operating-system, mtime/unix, mtime/universal, compression-level.
    Returns a dict, or None while the header is incomplete (input-underrun); header errors raise."""
    data = bytes(compressed[start:end])
    h = _ffi.GzipHeader()
    _ffi.check(_ffi.lib().tbz_gzip_header_parse(data, len(data), C.byref(h)))
    if h.verdict == 1:
        return None
    if h.verdict != 0:
        raise ThreeBzError(_ffi.verdict_name(h.verdict), h.verdict)

    def text(off, n):
        raw = data[off:off + n]
        try:
            return raw.decode("utf-8")                       # "rfc says 8859-1, but try utf8 anyway" (gzip.lisp:214-217)
        except UnicodeDecodeError:
            return raw.decode("iso8859-1")
    names = ((1, "text"), (2, "header-crc"), (4, "extra"), (8, "name"), (16, "comment"))
    return {
        "compression-method": "deflate",
        "flags": [n for bit, n in names if h.flags & bit],
        "extra": data[h.extra_off:h.extra_off + h.extra_len] if h.flags & 4 else None,
        "name": text(h.name_off, h.name_len) if h.flags & 8 else None,
        "comment": text(h.comment_off, h.comment_len) if h.flags & 16 else None,
        "operating-system": _GZ_OS[h.os] if h.os <= 13 else ("unknown", h.os),
        "mtime/unix": h.mtime or None,
        "mtime/universal": h.mtime + 2208988800 if h.mtime else None,   # (encode-universal-time 0 0 0 1 1 1970 0)
        "compression-level": {2: "maximum", 4: "fastest"}.get(h.xfl, h.xfl),
        "header-length": h.header_len,
    }


def decompress_gzip_members(compressed, output, max_members=1 << 20, ctx=None):
    """New entry point (the reference stops after the first member, gzip.lisp:279-286): decodes the
    concatenated gzip members of `compressed` back to back into `output`.  Returns
    (list of (out_len, in_used, crc32, verdict) per member, compressed bytes consumed)."""
    ctx = ctx or default_ctx()
    data = bytes(compressed)
    n = min(max_members, max(1, len(data) // 18 + 1))
    res = (_ffi.Result * n)()
    nm, used = C.c_uint64(), C.c_uint64()
    keep, addr = _addr(output)
    _ffi.check(ctx.L.tbz_inflate_gzip_members(ctx.h, data, len(data), addr, len(output), res, n, C.byref(nm), C.byref(used)), ctx.h)
    return [(r.out_len, r.in_used, r.checksum, r.verdict) for r in res[:nm.value]], used.value


def device_count():
    n = C.c_int32()
    check(lib().tbz_device_count(C.byref(n)))
    return n.value


def decompress_batch(members, format="zlib", capacities=None, ctx=None, flags=0, ctxs=None):
    """members: sequence of octet vectors; capacities: per-member output size (int or sequence).
    ctxs: several engine contexts (one per GPU): the members are partitioned over them on the host
    (tbz_inflate_batch_multi: one thread, context and stream per device, no collective).
    Returns a list of (buffer, count, verdict_code); a bad member never poisons the batch."""
    ctx = ctx or (ctxs[0] if ctxs else default_ctx())
    L = lib()
    n = len(members)
    if capacities is None:
        raise ValueError("capacities (one :output size per member) are required")
    caps = [capacities] * n if isinstance(capacities, int) else list(capacities)
    outs = [bytearray(c) for c in caps]
    marr = (_ffi.Member * max(1, n))()
    keep = []
    for i, (m, o) in enumerate(zip(members, outs)):
        k1, a1 = _addr(m)
        k2, a2 = _addr(o)
        keep += [k1, k2]
        marr[i] = _ffi.Member(a1, len(m), a2, len(o))
    rarr = (_ffi.Result * max(1, n))()
    if ctxs and len(ctxs) > 1:
        hs = (C.c_void_p * len(ctxs))(*[c.h for c in ctxs])
        check(L.tbz_inflate_batch_multi(hs, len(ctxs), fmt_code(format), marr, n, rarr, flags, None), ctx.h)
    else:
        check(L.tbz_inflate_batch(ctx.h, fmt_code(format), marr, n, rarr, flags, None), ctx.h)
    return [(outs[i], rarr[i].out_len, rarr[i].verdict) for i in range(n)]
