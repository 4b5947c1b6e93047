"""Runs the batched kernels under the CPU SIMT emulator (tests/emu): the same device source nvcc compiles,
executed on the host so that `-m "not gpu"` can check kernel logic against the oracle.  Test infrastructure
only — the product has no CPU path."""
import ctypes as C
import os
import subprocess

from threebz_b200 import _ffi

HERE = os.path.dirname(os.path.abspath(__file__))
EMU = os.path.join(HERE, "emu")
ROOT = os.path.dirname(HERE)
SO = os.path.join(EMU, "_emu.so")
_lib = None


def build(force=False):
    srcs = [os.path.join(EMU, f) for f in ("emu_driver.cpp", "cuda_emu.cpp")]
    csrc = os.path.join(ROOT, "3bz_b200", "csrc")
    deps = srcs + [os.path.join(EMU, "cuda_emu.h"), os.path.join(ROOT, "include", "threebz_cuda.h")] + \
        [os.path.join(csrc, f) for f in os.listdir(csrc) if f.endswith(".cuh")]
    if force or not os.path.exists(SO) or any(os.path.getmtime(d) > os.path.getmtime(SO) for d in deps):
        subprocess.check_call(["g++", "-O1", "-g", "-std=c++17", "-fPIC", "-shared", "-fsanitize=alignment",
                               "-fno-sanitize-recover=alignment", "-I", os.path.join(EMU, "shim"),
                               "-I", os.path.join(ROOT, "include"), "-I", csrc, "-o", SO] + srcs + ["-lpthread"])
    return SO


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.emu_inflate_batch.argtypes = [C.c_int, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint32, C.c_int, C.c_int]
        _lib.emu_inflate_batch.restype = C.c_int
        _lib.emu_find_block_start.argtypes = [C.c_char_p, C.c_uint64, C.c_uint32, C.c_uint32]
        _lib.emu_find_block_start.restype = C.c_uint32
        _lib.emu_split_adler.argtypes = [C.c_void_p, C.c_uint64]
        _lib.emu_split_adler.restype = C.c_uint32
        _lib.emu_tail_compose.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32]
        _lib.emu_tail_compose.restype = None
        _lib.emu_split_translate.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint32]
        _lib.emu_split_translate.restype = None
    return _lib


def run_batch(fmt, inputs, caps, flags=0, threads=None, in_mis=0, out_mis=0, variant=0):
    """Like tests.gpuutil.run_batch, on the emulator.  Returns (results, members that took the sequential kernel)."""
    L = lib()
    n = len(inputs)
    caps = [caps] * n if isinstance(caps, int) else list(caps)
    # every member 16-byte aligned (+ the requested misalignment), like the product's device arenas
    ioff, ooff, io, oo = [], [], 0, 0
    for d, c in zip(inputs, caps):
        ioff.append(io + in_mis); io += (len(d) + in_mis + 15 + 16) & ~15
        ooff.append(oo + out_mis); oo += (c + out_mis + 15 + 16) & ~15
    inbuf = (C.c_uint8 * (io + 64))()
    outbuf = (C.c_uint8 * (oo + 64))()
    ibase = (C.addressof(inbuf) + 15) & ~15
    obase = (C.addressof(outbuf) + 15) & ~15
    marr = (_ffi.Member * max(1, n))()
    for i, (d, c) in enumerate(zip(inputs, caps)):
        C.memmove(ibase + ioff[i], d, len(d))
        marr[i] = _ffi.Member(ibase + ioff[i], len(d), obase + ooff[i], c)
    rarr = (_ffi.Result * max(1, n))()
    nseq = L.emu_inflate_batch(_ffi.fmt_code(fmt), marr, n, rarr, flags, threads or (os.cpu_count() or 1), variant)
    res = []
    for i in range(n):
        r = rarr[i]
        res.append({"verdict": r.verdict, "out_len": r.out_len, "checksum": r.checksum, "where": r.where,
                    "in_used": r.in_used, "path": r.path,
                    "out": C.string_at(obase + ooff[i], r.out_len),
                    "raw": C.string_at(obase + ooff[i], caps[i])})     # the whole buffer (debugging a member that fell back)
    return res, nseq


def find_block_start(data, from_bit, to_bit):
    """tbzfast::find_block_start (the split decode's search) on the emulator: first dynamic-block start in bits
    [from_bit, to_bit) of the raw deflate stream `data`, or None."""
    r = lib().emu_find_block_start(data, len(data), from_bit, to_bit)
    return None if r == 0xffffffff else r


def _aligned(nbytes, mis=0):
    """(keep-alive buffer, address) of nbytes at a 64-byte aligned address + mis"""
    buf = (C.c_uint8 * (nbytes + 128))()
    return buf, ((C.addressof(buf) + 63) & ~63) + mis


def split_adler(data, mis=0):
    """Adler-32 through k_split_adler (the split decode's zlib checksum), data placed at a 64-byte aligned address + mis"""
    keep, a = _aligned(max(1, len(data)), mis)
    C.memmove(a, data, len(data))
    return lib().emu_split_adler(a, len(data))


def tail_compose(maps, stride):
    """k_tail_compose over a (nv, 32768) uint16 numpy array; returns the composed array"""
    import numpy as np
    nv = maps.shape[0]
    kin, ain = _aligned(maps.nbytes)
    kout, aout = _aligned(maps.nbytes)
    C.memmove(ain, maps.ctypes.data, maps.nbytes)
    lib().emu_tail_compose(ain, aout, nv, stride)
    return np.frombuffer(C.string_at(aout, maps.nbytes), dtype=np.uint16).reshape(maps.shape).copy()


def split_translate(offs, sym, out, grid=3, mis=0):
    """k_split_translate: offs (chunk output offsets + total), sym (uint16 numpy), out (bytearray holding the final
    tails); returns the translated bytes"""
    import numpy as np
    total = int(offs[-1])
    o = np.asarray(offs, dtype=np.uint64)
    ks, asym = _aligned(sym.nbytes + 64)
    C.memmove(asym, sym.ctypes.data, sym.nbytes)
    ko, aout = _aligned(total + 64, mis)
    C.memmove(aout, bytes(out), total)
    lib().emu_split_translate(o.ctypes.data, len(offs) - 1, asym, aout, total, grid)
    return C.string_at(aout, total)
