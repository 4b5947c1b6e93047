"""CPU: static checks of the Lisp shim (lisp/*.lisp).  No Common Lisp implementation exists in this image, so the shim
cannot be loaded here; these checks catch what a reader or compiler would reject first: unbalanced forms, bindings of
COMMON-LISP constants (`pi` bound as a variable), and C entry points the shim calls without a DEFCFUN
or that the header does not declare."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FILES = ["package.lisp", "ffi.lisp", "io.lisp", "api.lisp"]


def _strip(src):
    """source without comments and string contents (strings keep their quotes)"""
    out, i, n = [], 0, len(src)
    while i < n:
        c = src[i]
        if c == ";":
            while i < n and src[i] != "\n":
                i += 1
        elif c == '"':
            out.append('"')
            i += 1
            while i < n and src[i] != '"':
                i += 2 if src[i] == "\\" else 1
            out.append('"')
            i += 1
        elif c == "#" and i + 1 < n and src[i + 1] == "\\":
            out.append("#\\x")
            i += 3
        else:
            out.append(c)
            i += 1
    return "".join(out)


def _read(name):
    return open(os.path.join(ROOT, "lisp", name)).read()


def test_forms_are_balanced():
    for f in FILES:
        depth = 0
        for ch in _strip(_read(f)):
            depth += ch == "("
            depth -= ch == ")"
            assert depth >= 0, f
        assert depth == 0, f


def test_no_constant_is_bound():
    consts = {"pi", "t", "nil", "most-positive-fixnum"}
    for f in FILES:
        s = _strip(_read(f)).lower()
        # binding forms the shim uses: (with-pointer-to-vector-data (VAR ...)), (let ((VAR ...))), (lambda (VAR ...))
        for m in re.finditer(r"with-pointer-to-vector-data\s*\(\s*([^\s()]+)", s):
            assert m.group(1) not in consts, (f, m.group(0))
        for m in re.finditer(r"\(let\*?\s*\(\s*\(\s*([^\s()]+)", s):
            assert m.group(1) not in consts, (f, m.group(0))
        for m in re.finditer(r"\(lambda\s*\(([^()]*)\)", s):
            assert not (set(m.group(1).split()) & consts), (f, m.group(0))


def test_every_called_entry_point_is_bound_and_declared():
    ffi = _strip(_read("ffi.lisp"))
    bound = set(re.findall(r'defcfun\s+"(tbz_[a-z0-9_]+)"', _read("ffi.lisp")))
    header = open(os.path.join(ROOT, "include", "threebz_cuda.h")).read()
    declared = set(re.findall(r"\b(tbz_[a-z0-9_]+)\s*\(", header))
    assert bound <= declared, bound - declared
    called = set()
    for f in ("api.lisp", "io.lisp"):
        called |= set(re.findall(r"\((tbz-[a-z0-9-]+)", _strip(_read(f)).lower()))
    called -= {"tbz-member", "tbz-result", "tbz-verdict-name", "tbz-gzip-header"}
    lisp_names = {b.replace("_", "-") for b in bound}
    missing = {c for c in called if c not in lisp_names}
    assert not missing, missing
    assert "tbz_inflate_batch_multi" in bound and "tbz_session_consumed" in bound
    assert ffi  # (silence linters)
