"""ctypes binding of include/threebz_cuda.h (what the Lisp shim binds through CFFI in lisp/ffi.lisp).

No torch types cross this boundary.  Loading fails loudly if the CUDA extension is missing:
there is no CPU fallback behind this module.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.environ.get("TBZ_LIB") or os.path.join(HERE, "libthreebz_cuda.so")   # TBZ_LIB: tuning builds (tools/)

DEFLATE, ZLIB, GZIP = 0, 1, 2
FORMATS = {"deflate": DEFLATE, "zlib": ZLIB, "gzip": GZIP, ":deflate": DEFLATE, ":zlib": ZLIB, ":gzip": GZIP}
FINISHED, INPUT_UNDERRUN, OUTPUT_OVERFLOW = 0, 1, 2
AT_HEADER, AT_BODY, AT_TRAILER = 0, 1, 2
FLAG_DEVICE_PTRS, FLAG_NO_FASTPATH, FLAG_NO_SPLIT = 1, 2, 4
E_CUDA, E_NO_DEVICE, E_ARG, E_NOMEM, E_BUFFER_SWITCH, E_STATE = -1, -2, -3, -4, -5, -6

# every symbol include/threebz_cuda.h declares (tests check the .so exports all of them)
SYMBOLS = [
    "tbz_abi_version", "tbz_device_count", "tbz_ctx_create", "tbz_ctx_destroy", "tbz_strerror",
    "tbz_verdict_name", "tbz_ctx_last_error", "tbz_ctx_synchronize", "tbz_ctx_stream",
    "tbz_ctx_timer_start", "tbz_ctx_timer_stop", "tbz_ctx_kernel_timing", "tbz_ctx_last_kernel_ms",
    "tbz_ctx_launch_count",
    "tbz_host_alloc", "tbz_host_free", "tbz_host_register", "tbz_host_unregister",
    "tbz_device_alloc", "tbz_device_free", "tbz_memcpy_h2d", "tbz_memcpy_d2h",
    "tbz_inflate_batch", "tbz_inflate_single", "tbz_inflate_alloc", "tbz_free",
    "tbz_batch_prepare", "tbz_batch_launch", "tbz_batch_finish", "tbz_batch_destroy",
    "tbz_inflate_batch_multi", "tbz_partition",
    "tbz_session_create", "tbz_session_destroy", "tbz_session_set_output",
    "tbz_session_rebind_output", "tbz_session_replace_output", "tbz_session_decompress", "tbz_session_flags", "tbz_session_consumed",
    "tbz_gzip_header_parse", "tbz_inflate_gzip_members",
]


class Member(C.Structure):
    _fields_ = [("in_", C.c_void_p), ("in_len", C.c_uint64), ("out", C.c_void_p), ("out_cap", C.c_uint64)]


class Result(C.Structure):
    _fields_ = [("out_len", C.c_uint64), ("in_used", C.c_uint64), ("checksum", C.c_uint32),
                ("verdict", C.c_int32), ("where", C.c_uint32), ("path", C.c_uint32)]


class GzipHeader(C.Structure):
    _fields_ = [("verdict", C.c_int32), ("flags", C.c_uint32), ("mtime", C.c_uint32), ("xfl", C.c_uint32),
                ("os", C.c_uint32), ("header_crc", C.c_uint32),
                ("extra_off", C.c_uint64), ("extra_len", C.c_uint64), ("name_off", C.c_uint64), ("name_len", C.c_uint64),
                ("comment_off", C.c_uint64), ("comment_len", C.c_uint64), ("header_len", C.c_uint64)]


class EngineError(RuntimeError):
    def __init__(self, status, msg):
        super().__init__("threebz_cuda: %s (status %d)" % (msg, status))
        self.status = status


_lib = None


def lib():
    """Loads libthreebz_cuda.so (built in-tree by 3bz_b200/build.py).  Raises if it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO):
        raise ImportError("libthreebz_cuda.so is not built (run `python __graft_entry__.py` or "
                          "3bz_b200/build.py); this package has no CPU fallback")
    L = C.CDLL(SO)
    vp, u64, i32, u32 = C.c_void_p, C.c_uint64, C.c_int32, C.c_uint32
    P = C.POINTER
    sig = {
        "tbz_abi_version": (i32, []),
        "tbz_device_count": (i32, [P(i32)]),
        "tbz_ctx_create": (i32, [i32, u64, P(vp)]),
        "tbz_ctx_destroy": (i32, [vp]),
        "tbz_strerror": (C.c_char_p, [i32]),
        "tbz_verdict_name": (C.c_char_p, [i32]),
        "tbz_ctx_last_error": (C.c_char_p, [vp]),
        "tbz_ctx_synchronize": (i32, [vp]),
        "tbz_ctx_stream": (i32, [vp, P(vp)]),
        "tbz_ctx_timer_start": (i32, [vp]),
        "tbz_ctx_timer_stop": (i32, [vp, P(C.c_float)]),
        "tbz_ctx_kernel_timing": (i32, [vp, i32]),
        "tbz_ctx_last_kernel_ms": (i32, [vp, P(C.c_float)]),
        "tbz_ctx_launch_count": (i32, [vp, P(u64)]),
        "tbz_host_alloc": (i32, [u64, P(vp)]),
        "tbz_host_free": (i32, [vp]),
        "tbz_host_register": (i32, [vp, u64, u32]),
        "tbz_host_unregister": (i32, [vp]),
        "tbz_device_alloc": (i32, [vp, u64, P(vp)]),
        "tbz_device_free": (i32, [vp, vp]),
        "tbz_memcpy_h2d": (i32, [vp, vp, vp, u64]),
        "tbz_memcpy_d2h": (i32, [vp, vp, vp, u64]),
        "tbz_inflate_batch": (i32, [vp, i32, P(Member), u64, P(Result), u32, P(C.c_float)]),
        "tbz_inflate_single": (i32, [vp, i32, vp, u64, vp, u64, P(Result), u32, P(C.c_float)]),
        "tbz_inflate_alloc": (i32, [vp, i32, vp, u64, P(vp), P(Result)]),
        "tbz_free": (None, [vp]),
        "tbz_batch_prepare": (i32, [vp, i32, P(Member), u64, u32, P(vp)]),
        "tbz_batch_launch": (i32, [vp]),
        "tbz_batch_finish": (i32, [vp, P(Result)]),
        "tbz_batch_destroy": (i32, [vp]),
        "tbz_inflate_batch_multi": (i32, [P(vp), i32, i32, P(Member), u64, P(Result), u32, P(C.c_float)]),
        "tbz_partition": (i32, [P(u64), u64, i32, P(i32)]),
        "tbz_session_create": (i32, [vp, i32, P(vp)]),
        "tbz_session_destroy": (i32, [vp]),
        "tbz_session_set_output": (i32, [vp, vp, u64]),
        "tbz_session_rebind_output": (i32, [vp, vp]),
        "tbz_session_replace_output": (i32, [vp, vp, u64]),
        "tbz_session_decompress": (i32, [vp, vp, u64, P(C.c_int64), P(i32)]),
        "tbz_session_flags": (i32, [vp, P(i32), P(i32), P(i32)]),
        "tbz_session_consumed": (i32, [vp, P(u64)]),
        "tbz_gzip_header_parse": (i32, [vp, u64, P(GzipHeader)]),
        "tbz_inflate_gzip_members": (i32, [vp, vp, u64, vp, u64, P(Result), u64, P(u64), P(u64)]),
    }
    for name, (res, args) in sig.items():
        f = getattr(L, name)
        f.restype = res
        f.argtypes = args
    _lib = L
    return L


def check(status, ctx=None):
    if status != 0:
        L = lib()
        msg = L.tbz_strerror(status).decode()
        detail = L.tbz_ctx_last_error(ctx).decode() if ctx is not None else L.tbz_ctx_last_error(None).decode()
        raise EngineError(status, msg + (": " + detail if detail else ""))


def fmt_code(f):
    if isinstance(f, str):
        return FORMATS[f.lower()]
    return int(f)


def verdict_name(v):
    return lib().tbz_verdict_name(v).decode()
