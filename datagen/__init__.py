"""Deterministic synthetic inputs for tests and bench.  Not part of the product.

Text comes from datagen/textgen.c (splitmix64, 4096-word vocabulary, log-uniform ranks);
compression is the system libz through Python's zlib module — level 6, memLevel 8,
Z_DEFAULT_STRATEGY, windowBits 15 (zlib) / 31 (gzip) / -15 (raw).
"""
import ctypes as C
import os
import subprocess
import zlib
from concurrent.futures import ThreadPoolExecutor

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libtbzgen.so")
_lib = None

WBITS = {"deflate": -15, "zlib": 15, "gzip": 31}


def build(force=False):
    src = os.path.join(_HERE, "textgen.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-o", _SO, src])
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        _lib.tbzgen_text.argtypes = [C.c_void_p, C.c_size_t, C.c_uint64]
        _lib.tbzgen_random.argtypes = [C.c_void_p, C.c_size_t, C.c_uint64]
        _lib.tbzgen_init()
    return _lib


def text(size, seed):
    buf = C.create_string_buffer(size)
    lib().tbzgen_text(buf, size, seed)
    return buf.raw


def random_bytes(size, seed):
    buf = C.create_string_buffer(size)
    lib().tbzgen_random(buf, size, seed)
    return buf.raw


def compress(data, fmt="zlib", level=6, strategy=zlib.Z_DEFAULT_STRATEGY, mem_level=8):
    co = zlib.compressobj(level, zlib.DEFLATED, WBITS[fmt], mem_level, strategy)
    return co.compress(data) + co.flush()


def member(size, seed, fmt="zlib", level=6):
    """(plain, compressed) for text(size, seed)."""
    p = text(size, seed)
    return p, compress(p, fmt, level)


def members(n, size, seed0, fmt="zlib", level=6, threads=None):
    """n members with seeds seed0..seed0+n-1, compressed on `threads` host threads."""
    threads = threads or os.cpu_count() or 1
    lib()
    with ThreadPoolExecutor(threads) as ex:
        return list(ex.map(lambda i: member(size, seed0 + i, fmt, level), range(n)))
