// inflate_fast.cuh — the batched fast path: one CTA per member, every thread a decode lane.
//
// Inside one Huffman block the compressed bits are cut into NT equal sub-chunks.  Lane i starts
// decoding at the first bit of sub-chunk i *speculatively* (only lane 0 is known to start on a
// symbol boundary) and relies on the self-synchronisation of Huffman streams:
//   phase 1a  every lane decodes its sub-chunk into a token stream (global scratch, coalesced
//             rows) and marks its token-start bits in a bitmap (shared memory)
//   phase 1b  every lane keeps decoding past its sub-chunk end until it lands on a bit the next
//             lane(s) marked — from there on both decodes are identical (same tables, same bit,
//             same state), so the rest of that lane's tokens are proven correct
//   resolve   pointer-doubling over "who synchronised into whom" from lane 0 gives the set of
//             proven lanes, their entry points, token ranges and output sizes (prefix sum)
//   phase 2   every proven lane replays its tokens into a 64 KiB output ring in shared memory:
//             literals are byte stores, matches are copied in <=16-byte pieces once their source
//             range is final (per-lane progress words; overlapping matches copy through the
//             period); the ring is flushed to global memory with coalesced 16-byte stores and
//             the Adler-32 / CRC-32 is folded in from shared memory on the way out.
// Anything this kernel cannot prove clean (stored blocks, malformed codes, truncated input, too
// small an output buffer, checksum mismatch ...) is queued for the sequential kernel, which
// reproduces the reference's exact verdict.  Replaces deflate.lisp:465-509,673-702 (decode),
// :244-359 (copy-history), huffman-tree.lisp:99-218 (tables), checksums.lisp.
#pragma once
#include "tbz_device.cuh"

namespace tbzfast {

constexpr int NT = 256;                 // threads per CTA = decode lanes
constexpr int NWARP = NT / 32;
constexpr int KLL = 10, KD = 9;         // root table bits: lit/len, distance
constexpr uint32_t RING = 65536u, RMASK = RING - 1u;
constexpr int TOKCAP = 160;             // tokens a lane may emit per round (sub-chunk + overrun)
constexpr uint32_t S_MAX = 992, S_MIN = 256;   // sub-chunk size in bits (bitmap <= 32 KiB - 4)
constexpr uint32_t PIECE = 16;          // bytes copied per readiness check

constexpr uint32_t E_LONG = 0x00000300u, E_INVALID = 0x00010300u;   // table specials (code length 0)
constexpr uint32_t TOK_MATCH = 0x80000000u, TOK_EOB = 0x40000000u;

enum { ST_IDLE = 0, ST_END, ST_SYNC, ST_EOB, ST_CAP, ST_BAD };

struct Canon16 { uint16_t first[16], count[16], base[16]; uint16_t maxlen, nsyms; };

struct Smem {
  alignas(16) uint8_t ring[RING];        // output ring; its free half doubles as the sync bitmap
  uint32_t lut_ll[1 << KLL];
  uint32_t lut_d[1 << KD];
  uint32_t lut_cl[128];
  uint32_t crc_tab[256];
  Canon16 c_ll, c_d, c_cl;
  uint16_t sorted_ll[288], sorted_d[32], sorted_cl[32];
  uint8_t lens[352];                     // [0,19) code-length code, [32,352) lit/len + distance
  uint16_t run[2][16];                   // running offsets of the two table-building warps
  uint32_t e_pos[NT];                    // where the lane's decode stopped
  uint32_t entry[NT];                    // proven entry point of the lane (bit position)
  uint32_t outb[NT];                     // output bytes of the lane's proven range
  uint32_t obase[NT + 1];                // absolute output offset of each lane's range
  uint32_t prog[NT];                     // output position up to which the lane's bytes are final
  uint16_t nxt[2][NT + 1];               // successor lane (pointer doubling, double buffered)
  uint16_t ntok[NT], gtok[NT];
  uint8_t status[NT];
  uint8_t truth[NT + 1];
  uint16_t blk_owner[RING / 64];
  uint32_t wscan[NWARP];
  unsigned long long wsum[NWARP][2];
  // scalars
  uint32_t member;
  int fail;
  uint32_t term_pos; int term_status; uint32_t term_lane;
  uint32_t adler_s1, adler_s2, crc;
};

struct In {
  const uint32_t *w; uint32_t nwords; uint32_t pos0, end;
};
__device__ __forceinline__ uint32_t ldw(const In &in, uint32_t i) { return i < in.nwords ? __ldg(in.w + i) : 0u; }
__device__ __forceinline__ uint32_t peek32(const In &in, uint32_t pos) {
  uint32_t wi = pos >> 5;
  return __funnelshift_r(ldw(in, wi), ldw(in, wi + 1), pos & 31u);
}
__device__ __forceinline__ uint32_t byte_at(const In &in, uint32_t bytepos) {
  return (ldw(in, bytepos >> 2) >> (8 * (bytepos & 3))) & 0xff;
}

// ---- table entries ---------------------------------------------------------------------------
// lit/len: [3:0] code length, [6:4] extra bits, [9:8] kind (0 literal, 1 length, 2 end of block), [31:16] value
// dist   : [3:0] code length, [7:4] extra bits, [9:8] = 1, [31:16] base
__device__ __forceinline__ uint32_t ll_entry(uint32_t sym, uint32_t L) {
  if (sym < 256) return (sym << 16) | L;
  if (sym == 256) return (2u << 8) | L;
  if (sym > 285) return E_INVALID;                          // huffman-tree.lisp:176-177
  return ((uint32_t)c_len_base[sym - 257] << 16) | (1u << 8) | ((uint32_t)c_len_extra[sym - 257] << 4) | L;
}
__device__ __forceinline__ uint32_t d_entry(uint32_t sym, uint32_t L) {
  if (sym > 29) return E_INVALID;                           // huffman-tree.lisp:172-175
  return ((uint32_t)c_dist_base[sym] << 16) | (1u << 8) | ((uint32_t)c_dist_extra[sym] << 4) | L;
}

// canonical decode of the code that starts at bit 0 of `bits`, lengths lo..hi; returns (sym<<4)|L or 0
__device__ __forceinline__ uint32_t canon_lookup(const Canon16 &c, const uint16_t *sorted, uint32_t bits, int lo, int hi) {
  uint32_t rev = __brev(bits);
  for (int L = lo; L <= hi; L++) {
    uint32_t idx = (rev >> (32 - L)) - c.first[L];
    if (idx < c.count[L]) return ((uint32_t)sorted[c.base[L] + idx] << 4) | (uint32_t)L;
  }
  return 0;
}

// Warp-level canonical code construction from lens[0,n) (huffman-tree.lisp:107-183): length
// histogram with match_any groups, Kraft check, first code / base per length, symbols sorted by
// (length, symbol).  Returns 0 or a TBZ_ERR_* code (same order as the reference).
__device__ inline int warp_canon(const uint8_t *lens, int n, Canon16 &c, uint16_t *sorted, uint16_t *run, int lane) {
  if (lane < 16) run[lane] = 0;
  __syncwarp();
  for (int g = 0; g < n; g += 32) {
    int l = (g + lane < n) ? lens[g + lane] : 0;
    uint32_t m = __match_any_sync(TBZ_FULL, l);
    if (l && lane == __ffs(m) - 1) run[l] += (uint16_t)__popc(m);
    __syncwarp();
  }
  uint32_t cnt = (lane >= 1 && lane < 16) ? run[lane] : 0;
  __syncwarp();
  int err = 0, s = 1;
  uint32_t code = 0, b = 0, first = 0, base = 0;
#pragma unroll
  for (int L = 1; L <= 15; L++) {
    uint32_t cL = __shfl_sync(TBZ_FULL, cnt, L);
    if (!err) { s <<= 1; if ((int)cL > s) err = TBZ_ERR_OVERSUBSCRIBED; s -= (int)cL; }
    code <<= 1;
    if (lane == L) { first = code; base = b; }
    code += cL; b += cL;
  }
  uint32_t used = __ballot_sync(TBZ_FULL, cnt > 0);
  int maxlen = used ? 31 - __clz(used) : 0;
  if (lane < 16) { c.first[lane] = (uint16_t)first; c.count[lane] = (uint16_t)cnt; c.base[lane] = (uint16_t)base; run[lane] = (uint16_t)base; }
  if (lane == 0) { c.maxlen = (uint16_t)maxlen; c.nsyms = (uint16_t)b; }
  __syncwarp();
  if (err) return err;
  if (s > 0 && b > 1) return TBZ_ERR_INCOMPLETE;
  if (b == 1 && maxlen >= 11) return TBZ_ERR_TREE_TOO_LARGE;
  for (int g = 0; g < n; g += 32) {
    int l = (g + lane < n) ? lens[g + lane] : 0;
    uint32_t m = __match_any_sync(TBZ_FULL, l);
    uint32_t rank = __popc(m & ((1u << lane) - 1));
    if (l) sorted[run[l] + rank] = (uint16_t)(g + lane);
    __syncwarp();
    if (l && lane == __ffs(m) - 1) run[l] += (uint16_t)__popc(m);
    __syncwarp();
  }
  return 0;
}

// ---- per-lane bit reader: 64-bit buffer, 32-bit refills, next word prefetched ----------------
struct Bits { uint64_t bb; uint32_t bc, nw, wn; };
__device__ __forceinline__ void bits_init(Bits &b, const In &in, uint32_t pos) {
  uint32_t wi = pos >> 5, sh = pos & 31;
  uint64_t lo = ldw(in, wi), hi = ldw(in, wi + 1);
  b.bb = ((hi << 32) | lo) >> sh;
  b.bc = 64 - sh;
  b.nw = ldw(in, wi + 2);
  b.wn = wi + 3;
}
__device__ __forceinline__ void bits_refill(Bits &b, const In &in) {
  if (b.bc <= 32) {
    b.bb |= (uint64_t)b.nw << b.bc;
    b.bc += 32;
    b.nw = ldw(in, b.wn);
    b.wn++;
  }
}
__device__ __forceinline__ void bits_skip(Bits &b, uint32_t n) { b.bb >>= n; b.bc -= n; }

// One token.  Returns 0 literal, 1 match, 2 end of block, 3 invalid code.
__device__ __forceinline__ int decode_token(Bits &b, const In &in, const Smem &sm, uint32_t &tok, uint32_t &nbits, uint32_t &olen) {
  bits_refill(b, in);
  uint32_t e = sm.lut_ll[(uint32_t)b.bb & ((1u << KLL) - 1)];
  if ((e & 15) == 0) {
    if (e != E_LONG) return 3;
    uint32_t r = canon_lookup(sm.c_ll, sm.sorted_ll, (uint32_t)b.bb, KLL + 1, 15);
    if (!r) return 3;
    e = ll_entry(r >> 4, r & 15);
    if ((e & 15) == 0) return 3;
  }
  const uint32_t L = e & 15, kind = (e >> 8) & 3;
  if (kind == 0) { tok = e >> 16; nbits = L; olen = 1; bits_skip(b, L); return 0; }
  if (kind == 2) { tok = TOK_EOB; nbits = L; olen = 0; bits_skip(b, L); return 2; }
  const uint32_t xb = (e >> 4) & 7;
  const uint32_t len = (e >> 16) + ((uint32_t)(b.bb >> L) & ((1u << xb) - 1));
  bits_skip(b, L + xb);
  bits_refill(b, in);
  uint32_t d = sm.lut_d[(uint32_t)b.bb & ((1u << KD) - 1)];
  if ((d & 15) == 0) {
    if (d != E_LONG) return 3;
    uint32_t r = canon_lookup(sm.c_d, sm.sorted_d, (uint32_t)b.bb, KD + 1, 15);
    if (!r) return 3;
    d = d_entry(r >> 4, r & 15);
    if ((d & 15) == 0) return 3;
  }
  const uint32_t DL = d & 15, dxb = (d >> 4) & 15;
  const uint32_t dist = (d >> 16) + ((uint32_t)(b.bb >> DL) & ((1u << dxb) - 1));
  bits_skip(b, DL + dxb);
  tok = TOK_MATCH | ((dist - 1) << 8) | (len - 3);
  nbits = L + xb + DL + dxb;
  olen = len;
  return 1;
}

__device__ __forceinline__ uint32_t tok_outlen(uint32_t t) {
  return (t & TOK_MATCH) ? (t & 255u) + 3u : ((t & TOK_EOB) ? 0u : 1u);
}

// ---- CTA-wide helpers --------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cta_exclusive_scan(uint32_t v, Smem &sm, int tid, uint32_t &total) {
  const int lane = tid & 31, warp = tid >> 5;
  uint32_t x = v;
#pragma unroll
  for (int s = 1; s < 32; s <<= 1) {
    uint32_t y = __shfl_up_sync(TBZ_FULL, x, s);
    if (lane >= s) x += y;
  }
  if (lane == 31) sm.wscan[warp] = x;
  __syncthreads();
  uint32_t off = 0, tot = 0;
#pragma unroll
  for (int w = 0; w < NWARP; w++) { uint32_t t = sm.wscan[w]; if (w < warp) off += t; tot += t; }
  total = tot;
  __syncthreads();
  return off + x - v;
}

// Flush ring[a,b) (absolute output positions) to global memory, folding the bytes into the
// running Adler-32 (zlib) on the way.  Vector path when the member's output pointer is 16-byte
// aligned; bytewise otherwise.
__device__ inline void flush_range(Smem &sm, uint8_t *out, uint32_t a, uint32_t b, int fmt, int tid) {
  const uint32_t m = b - a;
  if (m == 0) return;
  unsigned long long sa = 0, sb = 0;   // sum d ; sum (m - j) d_j   (j relative to a)
  if ((((uintptr_t)out) & 15) == 0) {
    uint32_t head = (16 - (a & 15)) & 15;
    if (head > m) head = m;
    if ((uint32_t)tid < head) {
      uint32_t d = sm.ring[(a + tid) & RMASK];
      out[a + tid] = (uint8_t)d;
      sa += d; sb += (unsigned long long)(m - tid) * d;
    }
    const uint32_t body0 = a + head, nunits = (b - body0) >> 4;
    for (uint32_t u = tid; u < nunits; u += NT) {
      const uint32_t p = body0 + (u << 4);
      const uint4 v = *reinterpret_cast<const uint4 *>(&sm.ring[p & RMASK]);
      *reinterpret_cast<uint4 *>(out + p) = v;
      if (fmt == TBZ_ZLIB) {
        uint32_t s = __dp4a(v.x, 0x01010101u, 0u); s = __dp4a(v.y, 0x01010101u, s);
        s = __dp4a(v.z, 0x01010101u, s); s = __dp4a(v.w, 0x01010101u, s);
        uint32_t wj = __dp4a(v.x, 0x03020100u, 0u); wj = __dp4a(v.y, 0x07060504u, wj);
        wj = __dp4a(v.z, 0x0b0a0908u, wj); wj = __dp4a(v.w, 0x0f0e0d0cu, wj);
        sa += s;
        sb += (unsigned long long)(m - (p - a)) * s - wj;
      }
    }
    const uint32_t tail0 = body0 + (nunits << 4);
    if (tail0 + tid < b) {
      uint32_t d = sm.ring[(tail0 + tid) & RMASK];
      out[tail0 + tid] = (uint8_t)d;
      sa += d; sb += (unsigned long long)(b - (tail0 + tid)) * d;
    }
  } else {
    for (uint32_t p = a + tid; p < b; p += NT) {
      uint32_t d = sm.ring[p & RMASK];
      out[p] = (uint8_t)d;
      sa += d; sb += (unsigned long long)(b - p) * d;
    }
  }
  if (fmt == TBZ_ZLIB) {
    const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
    for (int s = 16; s; s >>= 1) {
      sa += __shfl_xor_sync(TBZ_FULL, sa, s);
      sb += __shfl_xor_sync(TBZ_FULL, sb, s);
    }
    if (lane == 0) { sm.wsum[warp][0] = sa; sm.wsum[warp][1] = sb; }
    __syncthreads();
    if (tid == 0) {
      unsigned long long A = 0, B = 0;
      for (int w = 0; w < NWARP; w++) { A += sm.wsum[w][0]; B += sm.wsum[w][1]; }
      // append m bytes: s2' = s2 + m*s1 + sum (m-j) d_j ; s1' = s1 + sum d
      unsigned long long s2 = (sm.adler_s2 + (unsigned long long)m % TBZ_ADLER_MOD * sm.adler_s1 + B % TBZ_ADLER_MOD) % TBZ_ADLER_MOD;
      sm.adler_s1 = (uint32_t)((sm.adler_s1 + A) % TBZ_ADLER_MOD);
      sm.adler_s2 = (uint32_t)s2;
    }
    __syncthreads();
  }
}

// CRC-32 of ring[a,b): every thread takes one contiguous slice, slices are merged pairwise with
// x^(8 len) shifts (the per-level shift is the square of the previous one).
__device__ inline void crc_range(Smem &sm, uint32_t a, uint32_t b, int tid) {
  const uint32_t m = b - a;
  if (m == 0) return;
  const uint32_t seg = (m + NT - 1) / NT;
  uint32_t lo = a + seg * tid, hi = lo + seg;
  if (lo > b) lo = b;
  if (hi > b) hi = b;
  uint32_t c = 0xffffffffu;
  for (uint32_t p = lo; p < hi; p++) c = (c >> 8) ^ sm.crc_tab[(c ^ sm.ring[p & RMASK]) & 0xff];
  c ^= 0xffffffffu;
  if (lo == hi) c = 0;
  // tree over NT slices; all full slices have length seg, trailing ones may be shorter or empty,
  // so each node carries its own length and uses the generic combine only when needed
  uint32_t len = hi - lo;
  uint32_t shift = crc_x8n(seg);                 // x^(8 seg), squared per level
  __shared__ uint32_t s_c[NT], s_l[NT];
  for (int s = 1; s < NT; s <<= 1) {
    s_c[tid] = c; s_l[tid] = len;
    __syncthreads();
    if ((tid & (2 * s - 1)) == 0 && tid + s < NT) {
      uint32_t oc = s_c[tid + s], ol = s_l[tid + s];
      if (ol) {
        uint32_t f = (ol == seg * (uint32_t)s) ? shift : crc_x8n(ol);
        c = crc_mulmod(f, c) ^ oc;
        len += ol;
      }
    }
    shift = crc_mulmod(shift, shift);
    __syncthreads();
  }
  if (tid == 0) sm.crc = crc_combine(sm.crc, c, m);   // crc of the empty prefix is 0
}

// ------------------------------------------------------------------------------------------------
// One member.  Returns true when the member was completed here, false when it must be redone by
// the sequential kernel.
// ------------------------------------------------------------------------------------------------
__device__ inline bool inflate_member(const DMember &mem, int fmt, tbz_result &res, Smem &sm,
                                      uint32_t *__restrict__ tokbuf, int tid) {
  const int lane = tid & 31, warp = tid >> 5;
  In in;
  {
    uintptr_t a = (uintptr_t)mem.in;
    uint32_t mis = (uint32_t)(a & 3);
    in.w = (const uint32_t *)(a - mis);
    in.pos0 = mis * 8;
    if (mem.in_len >= (1ull << 28)) return false;
    in.end = (mis + (uint32_t)mem.in_len) * 8;
    in.nwords = (in.end + 31) >> 5;
  }
  if (mem.out_cap >= (1ull << 31)) return false;
  uint32_t pos = in.pos0;
  // ---- wrapper header (zlib.lisp:108-126, gzip.lisp:113-177; optional gzip fields -> sequential kernel)
  if (fmt == TBZ_ZLIB) {
    if (in.end - pos < 16) return false;
    uint32_t cmf = byte_at(in, pos >> 3), flg = byte_at(in, (pos >> 3) + 1);
    if ((cmf * 256 + flg) % 31 || (cmf & 15) != 8 || (cmf >> 4) > 7 || (flg & 32)) return false;
    pos += 16;
  } else if (fmt == TBZ_GZIP) {
    if (in.end - pos < 80) return false;
    uint32_t bp = pos >> 3;
    if (byte_at(in, bp) != 0x1f || byte_at(in, bp + 1) != 0x8b || byte_at(in, bp + 2) != 8 || byte_at(in, bp + 3) != 0) return false;
    pos += 80;
  }
  if (tid == 0) { sm.fail = 0; sm.adler_s1 = 1; sm.adler_s2 = 0; sm.crc = 0; }
  uint32_t A = 0;              // output bytes produced and flushed so far
  bool last = false;
  __syncthreads();

  while (!last) {
    // ================= block header (deflate.lisp:518-528, :577-669) =================
    if (in.end - pos < 3) return false;
    const uint32_t hdr = peek32(in, pos) & 7;
    pos += 3;
    last = hdr & 1;
    const uint32_t btype = hdr >> 1;
    int hlit, hdist;
    if (btype == 1) {
      hlit = 288; hdist = 32;
      for (int i = tid; i < 320; i += NT) sm.lens[32 + i] = i < 144 ? 8 : i < 256 ? 9 : i < 280 ? 7 : i < 288 ? 8 : 5;
    } else if (btype == 2) {
      if (in.end - pos < 14) return false;
      const uint32_t v = peek32(in, pos);
      hlit = (v & 31) + 257; hdist = ((v >> 5) & 31) + 1;
      const int ncl = ((v >> 10) & 15) + 4;
      if (in.end - pos < 14u + 3u * ncl) return false;
      if (warp == 0) {
        if (lane < 19) sm.lens[lane] = 0;
        __syncwarp();
        if (lane < ncl) sm.lens[c_clen_order[lane]] = peek32(in, pos + 14 + 3 * lane) & 7;
        __syncwarp();
        int err = warp_canon(sm.lens, 19, sm.c_cl, sm.sorted_cl, sm.run[0], lane);
        if (!err && sm.c_cl.nsyms == 0) err = TBZ_ERR_INVALID_SYMBOL;
        if (!err) {
          for (int e = lane; e < 128; e += 32) {
            uint32_t r = canon_lookup(sm.c_cl, sm.sorted_cl, (uint32_t)e, 1, 7);
            sm.lut_cl[e] = r;               // (sym << 4) | L, 0 = no code
          }
          __syncwarp();
          if (lane == 0) {
            // the code lengths themselves: one lane, table driven (deflate.lisp:626-669)
            uint32_t p = pos + 14 + 3 * ncl;
            int idx = 0, lastlen = 0xff;
            const int total = hlit + hdist;
            while (idx < total) {
              uint32_t w = peek32(in, p);
              uint32_t r = sm.lut_cl[w & 127];
              if (!r) { err = TBZ_ERR_INVALID_SYMBOL; break; }
              int L = r & 15, sym = r >> 4;
              int xb = sym < 16 ? 0 : sym == 16 ? 2 : sym == 17 ? 3 : 7;
              if (p + L + xb > in.end) { err = TBZ_INPUT_UNDERRUN; break; }
              uint32_t extra = (w >> L) & ((1u << xb) - 1);
              p += L + xb;
              int rep, val;
              if (sym < 16) { rep = 1; val = sym; lastlen = sym; }
              else if (sym == 16) { if (lastlen >= 16) { err = TBZ_ERR_REPEAT_NO_PREV; break; } rep = 3 + extra; val = lastlen; }
              else { rep = (sym == 17 ? 3 : 11) + extra; val = 0; lastlen = 0; }
              if (idx + rep > total) { err = TBZ_ERR_REPEAT_OVERRUN; break; }
              for (int k = 0; k < rep; k++) sm.lens[32 + idx + k] = (uint8_t)val;
              idx += rep;
            }
            sm.term_pos = p;
          }
        }
        err = __shfl_sync(TBZ_FULL, err, 0) | err;
        if (err && lane == 0) sm.fail = 1;
      }
      __syncthreads();
      if (sm.fail) return false;
      pos = sm.term_pos;
    } else {
      return false;                        // stored / reserved block type: sequential kernel
    }
    __syncthreads();
    // ================= tables (huffman-tree.lisp:99-218) =================
    if (warp == 0) { if (warp_canon(sm.lens + 32, hlit, sm.c_ll, sm.sorted_ll, sm.run[0], lane) && lane == 0) sm.fail = 1; }
    else if (warp == 1) { if (warp_canon(sm.lens + 32 + hlit, hdist, sm.c_d, sm.sorted_d, sm.run[1], lane) && lane == 0) sm.fail = 1; }
    __syncthreads();
    if (sm.fail || sm.c_ll.nsyms == 0) return false;
    for (int e = tid; e < (1 << KLL); e += NT) {
      uint32_t r = canon_lookup(sm.c_ll, sm.sorted_ll, (uint32_t)e, 1, KLL);
      sm.lut_ll[e] = r ? ll_entry(r >> 4, r & 15) : (sm.c_ll.maxlen > KLL ? E_LONG : E_INVALID);
    }
    for (int e = tid; e < (1 << KD); e += NT) {
      uint32_t r = canon_lookup(sm.c_d, sm.sorted_d, (uint32_t)e, 1, KD);
      sm.lut_d[e] = r ? d_entry(r >> 4, r & 15) : (sm.c_d.maxlen > KD ? E_LONG : E_INVALID);
    }
    __syncthreads();

    // ================= rounds over the block's compressed bits =================
    bool block_done = false;
    while (!block_done) {
      // ---- geometry of this round
      const uint32_t P0 = pos;
      const uint32_t winbase = P0 & ~31u;
      uint32_t S = ((in.end - winbase + NT - 1) / NT + 31) & ~31u;
      if (S > S_MAX) S = S_MAX;
      if (S < S_MIN) S = S_MIN;
      const uint32_t winend = winbase + S * NT;
      const uint32_t bm0 = (A + 3) >> 2;                  // bitmap lives in the ring half not holding history
      uint32_t *ring32 = reinterpret_cast<uint32_t *>(sm.ring);
      const uint32_t bmwords = (S * NT) >> 5;
      for (uint32_t w = tid; w < bmwords; w += NT) ring32[(bm0 + w) & (RING / 4 - 1)] = 0;
      __syncthreads();

      // ---- phase 1a: speculative decode of the lane's sub-chunk
      const uint32_t cstart = winbase + S * tid, cend = cstart + S;
      uint32_t p = tid == 0 ? P0 : cstart;
      uint32_t k = 0, ob = 0;
      int st = ST_IDLE;
      Bits b;
      if (p < in.end) {
        bits_init(b, in, p);
        uint32_t curw = (p - winbase) >> 5, acc = 0;
        for (;;) {
          if (p >= cend) { st = ST_END; break; }
          if (k >= TOKCAP) { st = ST_CAP; break; }
          const uint32_t rel = p - winbase, wi = rel >> 5;
          if (wi != curw) { ring32[(bm0 + curw) & (RING / 4 - 1)] = acc; acc = 0; curw = wi; }
          acc |= 1u << (rel & 31);
          uint32_t tok, nb, ol;
          const int kind = decode_token(b, in, sm, tok, nb, ol);
          if (kind == 3 || p + nb > in.end) { st = ST_BAD; break; }
          tokbuf[k * NT + tid] = tok;
          k++; p += nb; ob += ol;
          if (kind == 2) { st = ST_EOB; break; }
        }
        ring32[(bm0 + curw) & (RING / 4 - 1)] = acc;
      }
      __syncthreads();
      // ---- phase 1b: run on until the decode lands on a marked bit of a later lane
      uint32_t nx = NT;
      if (st == ST_END) {
        for (;;) {
          if (p >= winend) break;                                 // round ends here, block continues
          const uint32_t rel = p - winbase;
          if ((ring32[(bm0 + (rel >> 5)) & (RING / 4 - 1)] >> (rel & 31)) & 1u) { st = ST_SYNC; nx = rel / S; break; }
          if (k >= TOKCAP) { st = ST_CAP; break; }
          uint32_t tok, nb, ol;
          const int kind = decode_token(b, in, sm, tok, nb, ol);
          if (kind == 3 || p + nb > in.end) { st = ST_BAD; break; }
          tokbuf[k * NT + tid] = tok;
          k++; p += nb; ob += ol;
          if (kind == 2) { st = ST_EOB; break; }
        }
      }
      sm.e_pos[tid] = p; sm.status[tid] = (uint8_t)st; sm.ntok[tid] = (uint16_t)k;
      sm.nxt[0][tid] = (uint16_t)nx;
      sm.truth[tid] = tid == 0;
      if (tid == 0) { sm.nxt[0][NT] = NT; sm.truth[NT] = 0; }
      __syncthreads();
      // ---- resolve: lanes reachable from lane 0 through "synchronised into" edges are proven
      {
        int cur = 0;
        for (int r = 0; r < 8; r++) {
          const uint16_t n1 = sm.nxt[cur][tid];
          if (sm.truth[tid]) sm.truth[n1] = 1;
          sm.nxt[cur ^ 1][tid] = sm.nxt[cur][n1];
          if (tid == 0) sm.nxt[cur ^ 1][NT] = NT;
          cur ^= 1;
          __syncthreads();
        }
      }
      const bool proven = sm.truth[tid];
      if (proven) {
        if (st == ST_SYNC) sm.entry[nx] = p;
        else { sm.term_lane = tid; sm.term_status = st; sm.term_pos = p; }
      }
      if (tid == 0) sm.entry[0] = P0;
      __syncthreads();
      if (sm.term_status == ST_BAD || sm.term_status == ST_IDLE) return false;
      // ---- tokens decoded before the entry point are garbage: count and size them
      uint32_t g = 0, gb = 0;
      if (proven && tid != 0) {
        const uint32_t r0 = cstart - winbase, r1 = sm.entry[tid] - winbase;   // r1 in [r0, r0 + S)
        for (uint32_t w = r0 >> 5; w <= (r1 >> 5); w++) {
          uint32_t bits = ring32[(bm0 + w) & (RING / 4 - 1)];
          if (w == (r1 >> 5)) bits &= (1u << (r1 & 31)) - 1u;
          g += __popc(bits);
        }
        for (uint32_t i = 0; i < g; i++) gb += tok_outlen(tokbuf[i * NT + tid]);
      }
      uint32_t total;
      const uint32_t myout = proven ? ob - gb : 0;
      const uint32_t off = cta_exclusive_scan(myout, sm, tid, total);
      if ((unsigned long long)A + total > mem.out_cap) return false;       // overflow: sequential kernel
      sm.obase[tid] = A + off;
      if (tid == 0) sm.obase[NT] = A + total;
      sm.prog[tid] = A + off;
      sm.gtok[tid] = (uint16_t)g;
      __syncthreads();

      // ---- phase 2: replay tokens into the ring
      const uint32_t R_end = A + total;
      uint32_t opos = A + off;
      const uint32_t oend = opos + myout;
      uint32_t kk = g, rem = 0, dist = 0;
      const uint32_t kend = k;
      while (A < R_end) {
        // the ring keeps [A - 32K, A) as history, so this step may write up to A + 64K - min(A, 32K)
        uint32_t lim = A + RING - (A < 32768u ? A : 32768u);
        if (lim > R_end) lim = R_end;
        // block-owner map for readiness checks inside [A, lim): who owns the first byte of each
        // 64-byte output block (for the block that straddles A: who owns byte A)
        if (myout && opos < oend) {
          for (uint32_t bl = (opos + 63) >> 6; (bl << 6) < oend && (bl << 6) < lim; bl++)
            sm.blk_owner[bl & (RING / 64 - 1)] = (uint16_t)tid;
          if (opos == A) sm.blk_owner[(A >> 6) & (RING / 64 - 1)] = (uint16_t)tid;
        }
        __syncthreads();
        bool busy = proven && opos < oend && opos < lim;
        while (__syncthreads_or(busy)) {
          // a few tokens per barrier round to amortise the barrier
          for (int it = 0; it < 8 && busy; it++) {
            if (rem == 0) {
              if (kk >= kend) { busy = false; break; }
              const uint32_t t = tokbuf[kk * NT + tid];
              kk++;
              if (t & TOK_MATCH) {
                rem = (t & 255u) + 3u; dist = ((t >> 8) & 0x7fffu) + 1u;
                if (dist > opos) { sm.fail = 1; busy = false; break; }   // deflate.lisp:343-345
              } else if (t & TOK_EOB) { busy = false; break; }
              else {
                sm.ring[opos & RMASK] = (uint8_t)t;
                opos++;
                __threadfence_block();
                *(volatile uint32_t *)&sm.prog[tid] = opos;
                if (opos >= lim) busy = false;
                continue;
              }
            }
            // a piece of the pending match
            uint32_t n = rem < PIECE ? rem : PIECE;
            if (n > lim - opos) n = lim - opos;
            const uint32_t s0 = opos - dist;
            uint32_t s1 = s0 + n;                    // source bytes [s0, s1) must be final ...
            if (s1 > opos) s1 = opos;                // ... an overlapping copy reads only behind itself
            bool ready = true;
            const uint32_t mybase = sm.obase[tid];
            const uint32_t f1 = s1 < mybase ? s1 : mybase;   // the part of the source other lanes write is [s0, f1)
            if (s0 < mybase && f1 > A) {
              // lane owning byte f1-1, then downwards over every lane that covers [s0, f1)
              uint32_t o = sm.blk_owner[((f1 - 1) >> 6) & (RING / 64 - 1)];
              while (sm.obase[o + 1] <= f1 - 1) o++;
              for (;;) {
                const uint32_t ob1 = sm.obase[o + 1];
                const uint32_t need = ob1 < f1 ? ob1 : f1;
                if (*(volatile uint32_t *)&sm.prog[o] < need) { ready = false; break; }
                const uint32_t ob0 = sm.obase[o];
                if (ob0 <= s0 || ob0 <= A) break;              // everything below A is final
                o--;
              }
            }
            if (!ready) break;
            for (uint32_t i = 0; i < n; i++) sm.ring[(opos + i) & RMASK] = sm.ring[(s0 + i) & RMASK];
            opos += n; rem -= n;
            __threadfence_block();
            *(volatile uint32_t *)&sm.prog[tid] = opos;
            if (opos >= lim) busy = false;
          }
          if (sm.fail) busy = false;
        }
        if (sm.fail) return false;
        // ---- flush [A, lim) and fold it into the checksum
        if (fmt == TBZ_GZIP) crc_range(sm, A, lim, tid);
        flush_range(sm, mem.out, A, lim, fmt, tid);
        A = lim;
        __syncthreads();
      }
      // ---- how did the round end?
      pos = sm.term_pos;
      if (sm.term_status == ST_EOB) block_done = true;
      __syncthreads();
    }
  }
  // ================= trailer (zlib.lisp:80-96, gzip.lisp:82-106) =================
  pos = (pos + 7) & ~7u;
  uint32_t ck = 0;
  if (fmt == TBZ_ZLIB) {
    if (in.end - pos < 32) return false;
    const uint32_t bp = pos >> 3;
    const uint32_t t = (byte_at(in, bp) << 24) | (byte_at(in, bp + 1) << 16) | (byte_at(in, bp + 2) << 8) | byte_at(in, bp + 3);
    ck = sm.adler_s1 | (sm.adler_s2 << 16);
    if (t != ck) return false;
    pos += 32;
  } else if (fmt == TBZ_GZIP) {
    if (in.end - pos < 64) return false;
    const uint32_t bp = pos >> 3;
    const uint32_t t = byte_at(in, bp) | (byte_at(in, bp + 1) << 8) | (byte_at(in, bp + 2) << 16) | (byte_at(in, bp + 3) << 24);
    ck = sm.crc;
    if (t != ck) return false;
    pos += 64;
  }
  if (tid == 0) {
    res.out_len = A;
    res.in_used = (pos - in.pos0 + 7) >> 3;
    res.checksum = ck;
    res.verdict = TBZ_FINISHED;
    res.where = TBZ_AT_BODY;
    res.path = 1;
  }
  return true;
}

}  // namespace tbzfast
