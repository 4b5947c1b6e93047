#!/usr/bin/env python3
"""Tiny pure-Python inflate tokenizer (RFC 1951) for kernel design statistics and CPU models of the
device algorithms.  Not part of the product and not the parity oracle.
tokens(raw_deflate_bytes) -> list of ints: literal = byte value (0..255); match = (length << 16) | distance, tagged with bit 31."""
import sys

LEN_BASE = [3,4,5,6,7,8,9,10,11,13,15,17,19,23,27,31,35,43,51,59,67,83,99,115,131,163,195,227,258]
LEN_XB = [0,0,0,0,0,0,0,0,1,1,1,1,2,2,2,2,3,3,3,3,4,4,4,4,5,5,5,5,0]
DIST_BASE = [1,2,3,4,5,7,9,13,17,25,33,49,65,97,129,193,257,385,513,769,1025,1537,2049,3073,4097,6145,8193,12289,16385,24577]
DIST_XB = [0,0,0,0,1,1,2,2,3,3,4,4,5,5,6,6,7,7,8,8,9,9,10,10,11,11,12,12,13,13]
ORDER = [16,17,18,0,8,7,9,6,10,5,11,4,12,3,13,2,14,1,15]
MATCH = 1 << 31


def _table(lens):
    """canonical code -> dict (length, code) -> symbol, plus max length"""
    cnt = [0] * 16
    for l in lens:
        cnt[l] += 1
    cnt[0] = 0
    nxt, code = [0] * 16, 0
    for b in range(1, 16):
        code = (code + cnt[b - 1]) << 1
        nxt[b] = code
    tab = {}
    for s, l in enumerate(lens):
        if l:
            tab[(l, nxt[l])] = s
            nxt[l] += 1
    return tab


class _Bits:
    def __init__(self, data):
        self.d, self.pos = data, 0

    def bits(self, n):
        v = 0
        for i in range(n):
            v |= ((self.d[self.pos >> 3] >> (self.pos & 7)) & 1) << i
            self.pos += 1
        return v

    def sym(self, tab):
        code, l = 0, 0
        while True:
            code = (code << 1) | self.bits(1)
            l += 1
            s = tab.get((l, code))
            if s is not None:
                return s
            if l > 15:
                raise ValueError("bad code")


def tokens(data):
    b = _Bits(data)
    out = []
    while True:
        last, bt = b.bits(1), b.bits(2)
        if bt == 0:
            b.pos = (b.pos + 7) & ~7
            n = b.bits(16); b.bits(16)
            for _ in range(n):
                out.append(b.bits(8))
        else:
            if bt == 1:
                ll = _table([8] * 144 + [9] * 112 + [7] * 24 + [8] * 8)
                dd = _table([5] * 32)
            else:
                hlit, hdist, hclen = b.bits(5) + 257, b.bits(5) + 1, b.bits(4) + 4
                cl = [0] * 19
                for i in range(hclen):
                    cl[ORDER[i]] = b.bits(3)
                ct = _table(cl)
                lens = []
                while len(lens) < hlit + hdist:
                    s = b.sym(ct)
                    if s < 16: lens.append(s)
                    elif s == 16: lens += [lens[-1]] * (3 + b.bits(2))
                    elif s == 17: lens += [0] * (3 + b.bits(3))
                    else: lens += [0] * (11 + b.bits(7))
                ll, dd = _table(lens[:hlit]), _table(lens[hlit:])
            while True:
                s = b.sym(ll)
                if s < 256: out.append(s)
                elif s == 256: break
                else:
                    l = LEN_BASE[s - 257] + b.bits(LEN_XB[s - 257])
                    ds = b.sym(dd)
                    d = DIST_BASE[ds] + b.bits(DIST_XB[ds])
                    out.append(MATCH | (l << 16) | d)
        if last:
            return out


def expand(toks):
    o = bytearray()
    for t in toks:
        if t & MATCH:
            l, d = (t >> 16) & 0x7fff, t & 0xffff
            for _ in range(l):
                o.append(o[-d])
        else:
            o.append(t)
    return bytes(o)


if __name__ == "__main__":
    import os, zlib
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import datagen
    p, c = datagen.member(65536, 1000, "deflate")
    tk = tokens(c)
    assert expand(tk) == p
    print(len(tk), "tokens ok")
