// inflate_seq.cuh — the sequential, verdict-exact inflate kernel: one warp per member.
//
// This is the engine's reference-faithful path: it walks a member exactly in stream order and
// reproduces 3bz's verdict rules including the bytes produced before an
// underrun / overflow / error.  The fast kernels hand any member they cannot prove clean to
// this one.  It is still warp-cooperative:
//   * Huffman decode is a warp-wide canonical-code match — lane L tests "is the next L-bit
//     prefix a code of length L" against its own (first code, count) pair; one ballot picks the
//     length (replaces the nested-table walk of deflate.lisp:465-501 / huffman-tree.lisp:186-218)
//   * code-length histogram / first-code assignment are lane-parallel (huffman-tree.lisp:107-142)
//   * LZ77 copies, stored blocks and both checksums are lane-parallel (deflate.lisp:244-359,
//     :532-573, checksums.lisp)
#pragma once
#include "tbz_device.cuh"

namespace tbzseq {

struct BitIn {
  const uint32_t *w;   // 4-byte aligned base at or below the member's first byte
  uint64_t pos;        // next unread bit, relative to w
  uint64_t end;        // one past the last valid bit, relative to w
  uint64_t nwords;     // words that may be touched
};

__device__ __forceinline__ uint32_t peek32(const BitIn &b, uint64_t pos) {
  uint64_t wi = pos >> 5;
  uint32_t lo = wi < b.nwords ? __ldg(b.w + wi) : 0u;
  uint32_t hi = wi + 1 < b.nwords ? __ldg(b.w + wi + 1) : 0u;
  return __funnelshift_r(lo, hi, (uint32_t)pos & 31u);
}
__device__ __forceinline__ uint64_t avail(const BitIn &b) { return b.end - b.pos; }
__device__ __forceinline__ uint32_t byte_at(const BitIn &b, uint64_t bytepos /* relative to w */) {
  return (__ldg(b.w + (bytepos >> 2)) >> (8 * (bytepos & 3))) & 0xff;
}

// per-lane slice of a canonical Huffman code: lane L (1..15) owns the codes of length L
struct Canon {
  uint32_t first, count, base;
  int maxlen;      // longest used length (uniform)
  int nsyms;       // coded symbols (uniform); 0 = empty tree
};

// Build from code lengths lens[0,n) (shared memory).  Returns 0 or a TBZ_ERR_* verdict, with the
// reference's error order (huffman-tree.lisp:112-122, then the node-array bound :208-216).
__device__ inline int canon_build(const uint8_t *lens, int n, Canon &c, uint16_t *symtab,
                                  uint16_t *offs /*16, shared*/, int lane) {
  uint32_t cnt = 0;
  for (int i = 0; i < n; i++) cnt += (lens[i] == lane);
  if (lane == 0 || lane > 15) cnt = 0;
  int err = 0;
  int s = 1;
  uint32_t code = 0, b = 0;
  c.first = 0; c.base = 0;
#pragma unroll
  for (int L = 1; L <= 15; L++) {
    uint32_t cL = __shfl_sync(TBZ_FULL, cnt, L);
    if (!err) {
      s <<= 1;
      if ((int)cL > s) err = TBZ_ERR_OVERSUBSCRIBED;
      s -= (int)cL;
    }
    code <<= 1;
    if (lane == L) { c.first = code; c.base = b; }
    code += cL; b += cL;
  }
  c.count = cnt;
  c.nsyms = (int)b;
  uint32_t used = __ballot_sync(TBZ_FULL, cnt > 0);
  c.maxlen = used ? 31 - __clz(used) : 0;
  if (err) return err;
  if (s > 0 && c.nsyms > 1) return TBZ_ERR_INCOMPLETE;
  if (c.nsyms == 1 && c.maxlen >= 11) return TBZ_ERR_TREE_TOO_LARGE;
  if (c.nsyms == 0) return 0;
  if (lane >= 1 && lane <= 15) offs[lane] = (uint16_t)c.base;
  __syncwarp();
  if (lane == 0)
    for (int i = 0; i < n; i++) {
      int l = lens[i];
      if (l) symtab[offs[l]++] = (uint16_t)i;
    }
  __syncwarp();
  return 0;
}

// Decode one symbol from the 32 peeked bits w.  Returns the symbol and its length L, or
// -1 (input underrun) / -2 (invalid code) following deflate.lisp:361-461.
__device__ __forceinline__ int canon_decode(uint32_t w, uint64_t av, const Canon &c,
                                            const uint16_t *symtab, int lane, int &L) {
  if (c.nsyms == 0) { L = 0; return -2; }                 // all-invalid table (huffman-tree.lisp:156-157)
  uint32_t rev = __brev(w);
  uint32_t code = rev >> ((32 - lane) & 31);
  uint32_t idx = code - c.first;
  bool hit = lane >= 1 && idx < c.count;
  uint32_t m = __ballot_sync(TBZ_FULL, hit);
  if (!m) { L = 0; return av >= (uint64_t)c.maxlen ? -2 : -1; }
  L = __ffs(m) - 1;
  if ((uint64_t)L > av) return -1;
  uint32_t si = __shfl_sync(TBZ_FULL, c.base + idx, L);
  return symtab[si];
}

struct WarpSmem {
  uint8_t lens[32 + 320];   // [0,19) code-length code lengths, [32,352) lit/len + distance lengths
  uint16_t sym_ll[288];
  uint16_t sym_d[32];
  uint16_t sym_cl[32];
  uint16_t offs[16];
};

struct Out {
  uint8_t *p; uint64_t pos, cap;
};

// LZ77 copy of n bytes at distance d (d <= pos): lanes stride the destination; overlapping
// copies read through the period (k mod d), so every source byte is already final.
__device__ __forceinline__ void lz_copy(Out &o, uint32_t n, uint32_t d, int lane) {
  uint8_t *dst = o.p + o.pos;
  const uint8_t *src = dst - d;
  __syncwarp();
  if (d >= n) {
    for (uint32_t k = lane; k < n; k += 32) dst[k] = src[k];
  } else {
    for (uint32_t k = lane; k < n; k += 32) dst[k] = src[k % d];
  }
  o.pos += n;
}

// What a session keeps between two `decompress` calls (deflate.lisp:4-62 keeps every register of its state machine;
// here a call resumes at the last block boundary): everything below is relative to the first octet of the stream.
struct Resume {
  unsigned long long blk_bit;    // bit offset of the next block header (valid once header_done)
  unsigned long long blk_out;    // output bytes produced by the blocks before it
  unsigned long long ck_pos;     // output bytes the running checksum covers
  uint32_t ck;                   // Adler-32 (s1 | s2 << 16) or finalised CRC-32 of out[0, ck_pos)
  uint32_t header_done;          // the wrapper header has been parsed and accepted
};

// One member, one warp.  fmt = TBZ_DEFLATE / TBZ_ZLIB / TBZ_GZIP.  rs (sessions only): resume at the last block
// boundary, leave the new one behind; the checksum then runs over the new output bytes only.
__device__ inline void inflate_member(const DMember &m, int fmt, tbz_result &res, WarpSmem &sm,
                                      const uint32_t *crc_tab, int lane, Resume *rs = nullptr) {
  BitIn in;
  {
    uintptr_t a = (uintptr_t)m.in;
    uint32_t mis = (uint32_t)(a & 3);
    in.w = (const uint32_t *)(a - mis);
    in.pos = (uint64_t)mis * 8;
    in.end = ((uint64_t)mis + m.in_len) * 8;
    in.nwords = (in.end + 31) >> 5;
  }
  const uint64_t pos0 = in.pos;
  Out out{m.out, 0, m.out_cap};
  int verdict = -1;
  uint32_t where = TBZ_AT_HEADER;
  uint32_t hcrc_state = 0xffffffffu;   // gzip FHCRC over the header bytes

  // ---------------- wrapper headers ----------------
  const bool resumed = rs && rs->header_done;
  if (resumed) {                                           // (a session that is past its header)
    in.pos = pos0 + rs->blk_bit;
    out.pos = rs->blk_out;
  } else if (fmt == TBZ_ZLIB) {                            // zlib.lisp:108-126, :14-37
    if (avail(in) < 16) verdict = TBZ_INPUT_UNDERRUN;
    else {
      uint32_t cmf = byte_at(in, in.pos >> 3), flg = byte_at(in, (in.pos >> 3) + 1);
      in.pos += 16;
      if ((cmf * 256 + flg) % 31) verdict = TBZ_ERR_ZLIB_FCHECK;
      else if ((cmf & 15) != 8) verdict = TBZ_ERR_ZLIB_METHOD;
      else if ((cmf >> 4) > 7) verdict = TBZ_ERR_ZLIB_WINDOW;
      else if (flg & 32) verdict = TBZ_ERR_ZLIB_DICT;
    }
  } else if (fmt == TBZ_GZIP) {                            // gzip.lisp:113-266
    uint64_t bp = in.pos >> 3;
    const uint64_t be = in.end >> 3;
    uint32_t flg = 0;
#define TBZ_HB(v) do { v = byte_at(in, bp); hcrc_state = (hcrc_state >> 8) ^ crc_tab[(hcrc_state ^ v) & 0xff]; bp++; } while (0)
    uint32_t b0, b1;
    do {
      if (be - bp < 2) { verdict = TBZ_INPUT_UNDERRUN; break; }
      TBZ_HB(b0); TBZ_HB(b1);
      if (b0 != 0x1f || b1 != 0x8b) { verdict = TBZ_ERR_GZIP_MAGIC; break; }
      if (be - bp < 2) { verdict = TBZ_INPUT_UNDERRUN; break; }
      TBZ_HB(b0); TBZ_HB(flg);
      if (b0 != 8) { verdict = TBZ_ERR_GZIP_METHOD; break; }
      if (flg >> 5) { verdict = TBZ_ERR_GZIP_RESERVED; break; }
      if (be - bp < 4) { verdict = TBZ_INPUT_UNDERRUN; break; }
      TBZ_HB(b0); TBZ_HB(b0); TBZ_HB(b0); TBZ_HB(b0);      // MTIME
      if (be - bp < 2) { verdict = TBZ_INPUT_UNDERRUN; break; }
      TBZ_HB(b0); TBZ_HB(b0);                              // XFL, OS
      if (flg & 4) {                                       // FEXTRA
        if (be - bp < 2) { verdict = TBZ_INPUT_UNDERRUN; break; }
        TBZ_HB(b0); TBZ_HB(b1);
        uint32_t xlen = b0 | (b1 << 8);
        for (uint32_t i = 0; i < xlen; i++) {
          if (be - bp < 1) { verdict = TBZ_INPUT_UNDERRUN; break; }
          TBZ_HB(b0);
        }
        if (verdict >= 0) break;
      }
      for (int f = 8; f <= 16; f <<= 1)                    // FNAME, FCOMMENT
        if (flg & f) {
          for (;;) {
            if (be - bp < 1) { verdict = TBZ_INPUT_UNDERRUN; break; }
            TBZ_HB(b0);
            if (!b0) break;
          }
          if (verdict >= 0) break;
        }
      if (verdict >= 0) break;
      if (flg & 2) {                                       // FHCRC
        if (be - bp < 2) { verdict = TBZ_INPUT_UNDERRUN; break; }
        uint32_t want = (hcrc_state ^ 0xffffffffu) & 0xffff;
        b0 = byte_at(in, bp); b1 = byte_at(in, bp + 1); bp += 2;
        if ((b0 | (b1 << 8)) != want) { verdict = TBZ_ERR_GZIP_HCRC; break; }
      }
    } while (0);
#undef TBZ_HB
    in.pos = bp << 3;
  }

  // ---------------- deflate blocks (deflate.lisp:516-726) ----------------
  Canon ll, dd, cl;
  if (verdict < 0) where = TBZ_AT_BODY;
  if (rs && !resumed && verdict < 0 && lane == 0) { rs->header_done = 1; rs->blk_bit = in.pos - pos0; rs->blk_out = 0; }
  while (verdict < 0) {
    // :start-of-block
    if (avail(in) < 3) { verdict = TBZ_INPUT_UNDERRUN; break; }
    uint32_t hdr = peek32(in, in.pos) & 7;
    in.pos += 3;
    const bool last = hdr & 1;
    const uint32_t btype = hdr >> 1;
    if (btype == 3) { verdict = TBZ_ERR_BLOCK_TYPE; break; }
    if (btype == 0) {
      // :uncompressed-block / :copy-block
      in.pos = (in.pos + 7) & ~7ull;
      if (avail(in) < 32) { verdict = TBZ_INPUT_UNDERRUN; break; }
      uint32_t v = peek32(in, in.pos);
      in.pos += 32;
      if ((v >> 16) != ((~v) & 0xffff)) { verdict = TBZ_ERR_STORED_LEN; break; }
      uint32_t len = v & 0xffff;
      uint64_t have = avail(in) >> 3, room = out.cap - out.pos;
      uint32_t n = len;
      if (n > have) n = (uint32_t)have;
      if (n > room) n = (uint32_t)room;
      uint64_t sb = in.pos >> 3;
      for (uint32_t k = lane; k < n; k += 32) out.p[out.pos + k] = (uint8_t)byte_at(in, sb + k);
      out.pos += n; in.pos += (uint64_t)n * 8;
      if (n < len) {   // per byte: a full output buffer is reported before an empty input
        verdict = (out.pos >= out.cap) ? TBZ_OUTPUT_OVERFLOW : TBZ_INPUT_UNDERRUN;
        break;
      }
    } else {
      if (btype == 1) {
        // fixed code (huffman-tree.lisp:89-97)
        for (int i = lane; i < 288; i += 32) sm.lens[i] = i < 144 ? 8 : i < 256 ? 9 : i < 280 ? 7 : 8;
        if (lane < 32) sm.lens[288 + lane] = 5;
        __syncwarp();
        canon_build(sm.lens, 288, ll, sm.sym_ll, sm.offs, lane);
        canon_build(sm.lens + 288, 32, dd, sm.sym_d, sm.offs, lane);
      } else {
        // :dynamic-huffman-block — 26 bits at once (deflate.lisp:577-595)
        if (avail(in) < 26) { verdict = TBZ_INPUT_UNDERRUN; break; }
        uint32_t v = peek32(in, in.pos);
        in.pos += 26;
        const int hlit = (v & 31) + 257, hdist = ((v >> 5) & 31) + 1, hclen = (v >> 10) & 15;
        if (lane < 19) sm.lens[lane] = 0;
        __syncwarp();
        if (lane == 0) {
          sm.lens[16] = (v >> 14) & 7; sm.lens[17] = (v >> 17) & 7;
          sm.lens[18] = (v >> 20) & 7; sm.lens[0] = (v >> 23) & 7;
        }
        // :dht-len-table — the remaining 3*hclen bits at once (deflate.lisp:597-624)
        if (avail(in) < (uint64_t)(3 * hclen)) { verdict = TBZ_INPUT_UNDERRUN; break; }
        __syncwarp();
        if (lane < hclen) sm.lens[c_clen_order[4 + lane]] = (peek32(in, in.pos + 3 * lane)) & 7;
        in.pos += 3 * hclen;
        __syncwarp();
        int e = canon_build(sm.lens, 19, cl, sm.sym_cl, sm.offs, lane);
        if (e) { verdict = e; break; }
        __syncwarp();
        // :dht-len-table-data (deflate.lisp:626-669)
        int idx = 0, lastlen = 0xff;
        const int total = hlit + hdist;
        while (idx < total) {
          uint32_t w = peek32(in, in.pos);
          uint64_t av = avail(in);
          int L;
          int sym = canon_decode(w, av, cl, sm.sym_cl, lane, L);
          if (sym < 0) { verdict = sym == -1 ? TBZ_INPUT_UNDERRUN : TBZ_ERR_INVALID_SYMBOL; break; }
          int xb = sym < 16 ? 0 : sym == 16 ? 2 : sym == 17 ? 3 : 7;
          if ((uint64_t)(L + xb) > av) { verdict = TBZ_INPUT_UNDERRUN; break; }
          uint32_t extra = (w >> L) & ((1u << xb) - 1);
          in.pos += L + xb;
          int rep, val;
          if (sym < 16) { rep = 1; val = sym; lastlen = sym; }
          else if (sym == 16) {
            if (lastlen >= 16) { verdict = TBZ_ERR_REPEAT_NO_PREV; break; }
            rep = 3 + extra; val = lastlen;
          } else { rep = (sym == 17 ? 3 : 11) + extra; val = 0; }
          if (idx + rep > total) { verdict = TBZ_ERR_REPEAT_OVERRUN; break; }
          if (sym >= 17) lastlen = 0;
          for (int k = lane; k < rep; k += 32) sm.lens[32 + idx + k] = (uint8_t)val;
          idx += rep;
        }
        if (verdict >= 0) break;
        __syncwarp();
        // build-trees* (huffman-tree.lisp:272-287): lit/len first, then distance
        e = canon_build(sm.lens + 32, hlit, ll, sm.sym_ll, sm.offs, lane);
        if (!e) e = canon_build(sm.lens + 32 + hlit, hdist, dd, sm.sym_d, sm.offs, lane);
        if (e) { verdict = e; break; }
      }
      // :decode-compressed-data (deflate.lisp:673-702)
      for (;;) {
        uint32_t w = peek32(in, in.pos);
        uint64_t av = avail(in);
        int L;
        int sym = canon_decode(w, av, ll, sm.sym_ll, lane, L);
        if (sym < 0) { verdict = sym == -1 ? TBZ_INPUT_UNDERRUN : TBZ_ERR_INVALID_SYMBOL; break; }
        if (sym < 256) {
          in.pos += L;
          if (out.pos >= out.cap) { verdict = TBZ_OUTPUT_OVERFLOW; break; }
          if (lane == 0) out.p[out.pos] = (uint8_t)sym;
          out.pos++;
        } else if (sym == 256) {
          in.pos += L;
          break;
        } else {
          if (sym > 285) { verdict = TBZ_ERR_INVALID_SYMBOL; break; }
          int xb = c_len_extra[sym - 257];
          if ((uint64_t)(L + xb) > av) { verdict = TBZ_INPUT_UNDERRUN; break; }
          uint32_t len = c_len_base[sym - 257] + ((w >> L) & ((1u << xb) - 1));
          uint64_t p2 = in.pos + L + xb;
          uint32_t w2 = peek32(in, p2);
          uint64_t av2 = in.end - p2;
          int DL;
          int ds = canon_decode(w2, av2, dd, sm.sym_d, lane, DL);
          if (ds < 0) { verdict = ds == -1 ? TBZ_INPUT_UNDERRUN : TBZ_ERR_INVALID_SYMBOL; break; }
          if (ds > 29) { verdict = TBZ_ERR_INVALID_SYMBOL; break; }
          int dxb = c_dist_extra[ds];
          if ((uint64_t)(DL + dxb) > av2) { verdict = TBZ_INPUT_UNDERRUN; break; }
          uint32_t dist = c_dist_base[ds] + ((w2 >> DL) & ((1u << dxb) - 1));
          in.pos = p2 + DL + dxb;
          if (dist > out.pos) { verdict = TBZ_ERR_DISTANCE_TOO_FAR; break; }
          uint64_t room = out.cap - out.pos;
          uint32_t n = len > room ? (uint32_t)room : len;
          lz_copy(out, n, dist, lane);
          if (n < len) { verdict = TBZ_OUTPUT_OVERFLOW; break; }
        }
      }
      if (verdict >= 0) break;
    }
    if (last) { verdict = TBZ_FINISHED; break; }
    if (rs && lane == 0) { rs->blk_bit = in.pos - pos0; rs->blk_out = out.pos; }     // a block is complete: the next call starts here
  }

  // ---------------- checksum + trailer ----------------
  __syncwarp();
  uint32_t ck = 0;
  if (rs) {                                                // a session: only the bytes that are new since the last call
    const unsigned long long from = rs->ck_pos, n = out.pos - from;
    ck = rs->ck;
    if (fmt == TBZ_ZLIB) ck = adler32_warp(out.p + from, n, lane, from ? (ck & 0xffffu) : 1u, from ? (ck >> 16) : 0u);
    else if (fmt == TBZ_GZIP) { const uint32_t c2 = crc32_warp(out.p + from, n, crc_tab, lane); ck = from ? (n ? crc_combine(ck, c2, n) : ck) : c2; }
    __syncwarp();
    if (lane == 0) { rs->ck = ck; rs->ck_pos = out.pos; }
  } else if (fmt == TBZ_ZLIB) ck = adler32_warp(out.p, out.pos, lane);
  else if (fmt == TBZ_GZIP) ck = crc32_warp(out.p, out.pos, crc_tab, lane);
  if (verdict == TBZ_FINISHED && fmt != TBZ_DEFLATE) {
    in.pos = (in.pos + 7) & ~7ull;                          // byte-align (zlib.lisp:139, gzip.lisp:273)
    where = TBZ_AT_TRAILER;
    uint64_t bp = in.pos >> 3;
    if (avail(in) < 32) verdict = TBZ_INPUT_UNDERRUN;
    else if (fmt == TBZ_ZLIB) {                             // big-endian Adler-32 (zlib.lisp:80-96)
      uint32_t t = (byte_at(in, bp) << 24) | (byte_at(in, bp + 1) << 16) | (byte_at(in, bp + 2) << 8) | byte_at(in, bp + 3);
      in.pos += 32;
      if (t != ck) verdict = TBZ_ERR_CHECKSUM;
    } else {                                                // little-endian CRC-32, then ISIZE unchecked (gzip.lisp:82-106)
      uint32_t t = byte_at(in, bp) | (byte_at(in, bp + 1) << 8) | (byte_at(in, bp + 2) << 16) | (byte_at(in, bp + 3) << 24);
      in.pos += 32;
      if (t != ck) verdict = TBZ_ERR_CHECKSUM;
      else if (avail(in) < 32) verdict = TBZ_INPUT_UNDERRUN;
      else in.pos += 32;
    }
  }
  if (lane == 0) {
    res.out_len = out.pos;
    res.in_used = (in.pos - pos0 + 7) >> 3;
    res.checksum = ck;
    res.verdict = verdict;
    res.where = verdict == TBZ_INPUT_UNDERRUN ? where : TBZ_AT_BODY;
    res.path = 0;
  }
}

}  // namespace tbzseq
