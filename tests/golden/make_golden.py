#!/usr/bin/env python3
"""Regenerates tests/golden/* from the reference's OWN test data (run in the build container,
where /root/reference exists; the GPU box only sees the committed outputs).

* test.deflated            — byte copy of the reference fixture (data, not source):
                             8-byte LE size + raw deflate  (test-chunked-input.lisp:7-25)
* nayuki_vectors.json      — the 37 bit-string vectors of deflate-test.lisp:69-302, packed
                             LSB-first exactly as deflate-test does (deflate-test.lisp:38-43),
                             with the expected output or the 'eof / 'format marker.
"""
import json, os, re, shutil, hashlib, zlib

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def tokenize(src):
    """Minimal s-expression reader: returns nested lists; strings as ('str', s); drops #++ forms."""
    i, n = 0, len(src)

    def skip_ws():
        nonlocal i
        while i < n:
            if src[i].isspace():
                i += 1
            elif src[i] == ';':
                while i < n and src[i] != '\n':
                    i += 1
            else:
                break

    def read():
        nonlocal i
        skip_ws()
        if i >= n:
            return None
        c = src[i]
        if c == '(':
            i += 1
            out = []
            while True:
                skip_ws()
                if src[i] == ')':
                    i += 1
                    return out
                out.append(read())
        if c == '"':
            j = i + 1
            while src[j] != '"':
                j += 2 if src[j] == '\\' else 1
            s = src[i + 1:j]
            i = j + 1
            return ('str', s)
        if src.startswith('#++', i):
            i += 3
            read()           # discard the disabled form
            return read()
        if c == "'":
            i += 1
            return ['quote', read()]
        if src.startswith('#\\', i):
            j = i + 2
            while j < n and not src[j].isspace() and src[j] not in '()':
                j += 1
            tok = src[i:j]
            i = j
            return tok
        j = i
        while j < n and not src[j].isspace() and src[j] not in '()':
            j += 1
        tok = src[i:j]
        i = j
        return tok

    forms = []
    while True:
        f = read()
        if f is None:
            break
        forms.append(f)
    return forms


def pack(bits):
    bits = bits.replace(' ', '')
    out = bytearray((len(bits) + 7) // 8)
    for x, c in enumerate(bits):
        if c == '1':
            out[x // 8] |= 1 << (x % 8)
    return bytes(out), len(bits)


def ev(x, env):
    if isinstance(x, tuple):
        return x[1]
    if isinstance(x, str):
        return env[x.lower()]
    if isinstance(x, list) and x and x[0] == 'concatenate':
        return ''.join(ev(a, env) for a in x[2:])
    raise ValueError(x)


def collect(form, env, out, line_of):
    if not isinstance(form, list) or not form:
        return
    head = form[0]
    if head == 'deflate-test':
        bits = ev(form[1], env)
        exp = ev(form[2], env)
        marker = None
        if len(form) > 3:
            marker = form[3][1]
        out.append((bits, exp, marker))
    elif head == 'let':
        env = dict(env)
        for b in form[1]:
            env[b[0].lower()] = ev(b[1], env)
        for f in form[2:]:
            collect(f, env, out, line_of)


def main():
    shutil.copyfile(os.path.join(REF, "test.deflated"), os.path.join(HERE, "test.deflated"))
    raw = open(os.path.join(REF, "test.deflated"), "rb").read()
    size = int.from_bytes(raw[:8], "little")
    plain = zlib.decompress(raw[8:], -15)
    assert len(plain) == size
    meta = {"file": "test.deflated", "size_field": size, "payload_offset": 8,
            "sha256": hashlib.sha256(plain).hexdigest(),
            "crc32": "%08x" % zlib.crc32(plain), "adler32": "%08x" % zlib.adler32(plain),
            "note": "sha256/crc/adler of the libz inflation of bytes 8.. ; SURVEY.md section 4"}
    json.dump(meta, open(os.path.join(HERE, "test_deflated.json"), "w"), indent=1)

    src = open(os.path.join(REF, "deflate-test.lisp")).read()
    # line numbers of each top-level (deflate-test / (let form, in file order
    lines = [m.start() for m in re.finditer(r'^\((deflate-test|let) ', src, re.M)]
    line_no = [src.count('\n', 0, p) + 1 for p in lines]
    forms = [f for f in tokenize(src) if isinstance(f, list) and f and f[0] in ('deflate-test', 'let')]
    # the first (let* ...) scratch form is #++-disabled and starts with let*, so it is not matched
    assert len(forms) == len(line_no) == 37, (len(forms), len(line_no))
    vecs = []
    for f, ln in zip(forms, line_no):
        got = []
        collect(f, {}, got, None)
        assert len(got) == 1
        bits, exp, marker = got[0]
        data, nbits = pack(bits)
        vecs.append({"line": ln, "bits": bits.replace(' ', ''), "nbits": nbits,
                     "input_hex": data.hex(), "expected_hex": exp.replace(' ', '').lower(),
                     "marker": marker})
    json.dump(vecs, open(os.path.join(HERE, "nayuki_vectors.json"), "w"), indent=1)
    print("wrote", len(vecs), "vectors;", sum(v["marker"] is None for v in vecs), "known-answer")


if __name__ == "__main__":
    main()
