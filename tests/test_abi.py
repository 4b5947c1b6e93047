"""CPU: the C-ABI library loads, exports every symbol include/threebz_cuda.h declares, the host-side
logic works, and compute entry points fail loudly (no CPU fallback) when no device exists."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_symbols_exported(engine):
    from threebz_b200 import _ffi
    L = _ffi.lib()
    hdr = open(os.path.join(ROOT, "include", "threebz_cuda.h")).read()
    declared = set(re.findall(r"\b(tbz_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    assert declared == set(_ffi.SYMBOLS), declared ^ set(_ffi.SYMBOLS)
    for s in declared:
        assert hasattr(L, s), s
    assert L.tbz_abi_version() == 1


def test_struct_layout_matches_header(engine):
    from threebz_b200 import _ffi
    assert C.sizeof(_ffi.Member) == 32 and C.sizeof(_ffi.Result) == 32


def test_verdict_numbering_shared_with_oracle(engine, oracle):
    from threebz_b200 import _ffi
    hdr = open(os.path.join(ROOT, "include", "threebz_cuda.h")).read()
    ohdr = open(os.path.join(ROOT, "oracle", "oracle3bz.h")).read()
    a = dict(re.findall(r"TBZ_(ERR_[A-Z_]+) = (\d+)", hdr))
    b = dict(re.findall(r"O3_(ERR_[A-Z_]+) = (\d+)", ohdr))
    for k, v in a.items():
        assert b[k] == v, k
    assert _ffi.verdict_name(0) == "finished"


def test_partition_is_balanced(engine):
    from threebz_b200 import _ffi
    L = _ffi.lib()
    n, g = 1000, 8
    lens = (C.c_uint64 * n)(*[(i * 7919) % 5000 + 1 for i in range(n)])
    owner = (C.c_int32 * n)()
    assert L.tbz_partition(lens, n, g, owner) == 0
    load = [0] * g
    for i in range(n):
        assert 0 <= owner[i] < g
        load[owner[i]] += lens[i]
    assert max(load) - min(load) <= 5000
    assert L.tbz_partition(lens, n, 0, owner) == _ffi.E_ARG
    # equal-sized members: every device gets the same count
    lens = (C.c_uint64 * 64)(*[4096] * 64)
    owner = (C.c_int32 * 64)()
    L.tbz_partition(lens, 64, 4, owner)
    assert sorted(list(owner).count(d) for d in range(4)) == [16] * 4


def test_no_device_fails_loudly(engine):
    """Without a GPU the engine refuses to run; it never decodes on the CPU."""
    from threebz_b200 import _ffi
    L = _ffi.lib()
    n = C.c_int32()
    L.tbz_device_count(C.byref(n))
    if n.value > 0:
        pytest.skip("a CUDA device is present")
    h = C.c_void_p()
    assert L.tbz_ctx_create(0, 0, C.byref(h)) == _ffi.E_NO_DEVICE
    with pytest.raises(_ffi.EngineError):
        engine.decompress_vector(b"\x78\x9c\x03\x00\x00\x00\x00\x01", "zlib")


def test_product_does_not_import_oracle():
    """The oracle is test infrastructure: nothing under 3bz_b200/ may reference it."""
    pkg = os.path.join(ROOT, "3bz_b200")
    for dp, dn, fn in os.walk(pkg):
        for f in fn:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(dp, f), errors="replace").read()
                assert "oracle" not in src.lower().replace("no cpu oracle", ""), os.path.join(dp, f)
