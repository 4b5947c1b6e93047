"""Shared test inputs: the reference's fixtures and the config-5 edge mix."""
import json
import os
import zlib

import datagen

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden")
FMTS = ("deflate", "zlib", "gzip")


def test_deflated():
    raw = open(os.path.join(GOLD, "test.deflated"), "rb").read()
    meta = json.load(open(os.path.join(GOLD, "test_deflated.json")))
    return raw, meta


def nayuki():
    return json.load(open(os.path.join(GOLD, "nayuki_vectors.json")))


def splitmix(seed):
    s = seed & (2**64 - 1)
    while True:
        s = (s + 0x9E3779B97F4A7C15) & (2**64 - 1)
        z = s
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & (2**64 - 1)
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & (2**64 - 1)
        yield z ^ (z >> 31)


def rle_runs():
    out = bytearray()
    for i in range(600):
        out += bytes([i % 251]) * (1 + (37 * i) % 700)
    return bytes(out)


def edge_plain():
    """name -> (plain bytes, compress kwargs)"""
    text = datagen.text(200000, 77)
    return {
        "stored": (text, dict(level=0)),
        "incompressible": (datagen.random_bytes(200000, 78), dict(level=6)),
        "fixed": (text[:60000], dict(level=6, strategy=zlib.Z_FIXED)),
        "huffman_only": (text[:60000], dict(level=6, strategy=zlib.Z_HUFFMAN_ONLY)),
        "zeros": (bytes(200000), dict(level=6)),
        "rle": (rle_runs(), dict(level=6, strategy=zlib.Z_RLE)),
        "period3": (b"abc" * 70000, dict(level=6)),
        "empty": (b"", dict(level=6)),
        "one_byte": (b"x", dict(level=6)),
        "text64k": (datagen.text(65536, 1000), dict(level=6)),
        "text_l1": (text[:100000], dict(level=1)),
        "text_l9": (text[:100000], dict(level=9)),
        "text300k": (datagen.text(300000, 79), dict(level=6)),
    }


_cache = {}


def edge_streams():
    """list of (name, fmt, compressed, plain)"""
    if "edge" not in _cache:
        out = []
        for name, (plain, kw) in edge_plain().items():
            for fmt in FMTS:
                out.append((name, fmt, datagen.compress(plain, fmt, **kw), plain))
        _cache["edge"] = out
    return _cache["edge"]


def gzip_with_header_fields(plain, extra=b"EXTRA!", name=b"file.txt", comment=b"a comment", hcrc=True):
    """gzip member with FEXTRA/FNAME/FCOMMENT/FHCRC set (libz never writes these by default)."""
    flg = (4 if extra is not None else 0) | (8 if name is not None else 0) | (16 if comment is not None else 0) | (2 if hcrc else 0)
    hdr = bytearray(b"\x1f\x8b\x08" + bytes([flg]) + b"\x01\x02\x03\x04" + b"\x00\x03")
    if extra is not None:
        hdr += len(extra).to_bytes(2, "little") + extra
    if name is not None:
        hdr += name + b"\0"
    if comment is not None:
        hdr += comment + b"\0"
    if hcrc:
        hdr += (zlib.crc32(bytes(hdr)) & 0xffff).to_bytes(2, "little")
    body = datagen.compress(plain, "deflate")
    return bytes(hdr) + body + zlib.crc32(plain).to_bytes(4, "little") + (len(plain) & 0xffffffff).to_bytes(4, "little")
