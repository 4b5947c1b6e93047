// inflate_decode.cuh — phase one of the batched fast path: Huffman decode into a token stream.
// One CTA per member, every thread a decode lane.
//
// Inside one Huffman block the compressed bits are cut into NT equal sub-chunks.  Lane i starts
// decoding at the first bit of sub-chunk i *speculatively* (only lane 0 is known to start on a
// symbol boundary) and relies on the self-synchronisation of Huffman streams:
//   1a  every lane decodes its sub-chunk into its own token list (a slab in global memory,
//       16-byte stores) and marks its token-start bits in a bitmap (shared memory)
//   1b  every lane keeps decoding past its sub-chunk end until it lands on a bit a later lane
//       marked — from there on both decodes are identical (same tables, same bit, same state),
//       so the rest of that lane's list is proven correct
//   1c  pointer doubling over "who synchronised into whom" from lane 0 gives the proven lanes,
//       their entry points (tokens before an entry point are dropped) and the round's output size
// Phase two (inflate_resolve.cuh) turns the token stream into bytes.  Anything this kernel cannot
// prove clean (stored blocks, malformed codes, truncated input, too small an output buffer ...)
// is queued for the sequential kernel, which reproduces the reference's exact verdict.
// Replaces deflate.lisp:465-509,673-702 (decode) and huffman-tree.lisp:99-218 (tables).
#pragma once
#include "tbz_device.cuh"

namespace tbzfast {

constexpr int NT = 256;                 // threads per CTA = decode lanes
constexpr int NWARP = NT / 32;
constexpr int KLL = 10, KD = 9;         // root table bits: lit/len, distance
constexpr int TOKCAP = 160;             // tokens a lane may emit per round (sub-chunk + overrun)
constexpr uint32_t S_MAX = 992, S_MIN = 256;   // sub-chunk size in bits
constexpr uint32_t SYNC_SLACK = 640;           // bits a lane searches for its sync point before the barrier
constexpr uint32_t BMWORDS = S_MAX * NT / 32;  // sync bitmap, one bit per compressed bit of the round

// A slab holds the token lists of one round: header, then NT lists of TOKCAP tokens.
struct SlabHdr {
  uint32_t next;        // next slab of the member, or NO_SLAB
  uint32_t out_bytes;   // output bytes of the round
  uint32_t ntokens;     // proven tokens of the round
  uint32_t pad;
  uint32_t gn[NT];      // per lane: first proven token | (end << 16); 0 = lane not proven
  uint32_t tb[NT];      // per lane: number of proven tokens in the lanes before it (flat token index base)
};
constexpr uint32_t SLAB_WORDS = sizeof(SlabHdr) / 4 + NT * TOKCAP;
constexpr uint32_t NO_SLAB = 0xffffffffu;

// what phase one leaves per member for phase two
struct P1Rec {
  uint32_t status;      // 1 = token stream complete, 0 = member goes to the sequential kernel
  uint32_t first_slab;
  uint32_t out_len;     // total output bytes
  uint32_t end_pos;     // bit position (relative to the 4-byte aligned input base) after the last block
};

constexpr uint32_t E_LONG = 0x00000300u, E_INVALID = 0x00010300u;   // table specials (code length 0)
constexpr uint32_t TOK_MATCH = 0x80000000u, TOK_EOB = 0x40000000u;

enum { ST_IDLE = 0, ST_END, ST_SYNC, ST_EOB, ST_CAP, ST_BAD };

struct Canon16 { uint16_t first[16], count[16], base[16]; uint16_t maxlen, nsyms; };

struct Smem {
  uint32_t bitmap[BMWORDS];              // token-start bits of the current round
  uint32_t lut_ll[1 << KLL];
  uint32_t lut_d[1 << KD];
  uint32_t lut_cl[128];
  Canon16 c_ll, c_d, c_cl;
  uint16_t sorted_ll[288], sorted_d[32], sorted_cl[32];
  uint8_t lens[352];                     // [0,19) code-length code, [32,352) lit/len + distance
  uint16_t run[2][16];                   // running offsets of the two table-building warps
  uint32_t entry[NT];                    // proven entry point of the lane (bit position)
  uint16_t nxt[2][NT + 1];               // successor lane (pointer doubling, double buffered)
  uint8_t truth[NT + 1];
  uint32_t wscan[NWARP], wscan2[NWARP];
  // scalars
  uint32_t member;
  int fail;
  uint32_t term_pos; int term_status;
  uint32_t slab;
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
__device__ __forceinline__ void cta_sum2(uint32_t a, uint32_t b, Smem &sm, int tid, uint32_t &ta, uint32_t &tb) {
  const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
  for (int s = 16; s; s >>= 1) { a += __shfl_xor_sync(TBZ_FULL, a, s); b += __shfl_xor_sync(TBZ_FULL, b, s); }
  if (lane == 0) { sm.wscan[warp] = a; sm.wscan2[warp] = b; }
  __syncthreads();
  ta = 0; tb = 0;
#pragma unroll
  for (int w = 0; w < NWARP; w++) { ta += sm.wscan[w]; tb += sm.wscan2[w]; }
  __syncthreads();
}

// token list writer: four tokens per 16-byte store
struct TokW {
  uint32_t *list; uint32_t q0, q1, q2, q3;
  __device__ __forceinline__ void push(uint32_t tok, uint32_t k) {
    q0 = q1; q1 = q2; q2 = q3; q3 = tok;
    if ((k & 3) == 3) *reinterpret_cast<uint4 *>(list + (k & ~3u)) = make_uint4(q0, q1, q2, q3);
  }
  __device__ __forceinline__ void finish(uint32_t k) {     // k tokens pushed so far
    const uint32_t r = k & 3, b = k & ~3u;
    if (r == 1) list[b] = q3;
    else if (r == 2) { list[b] = q2; list[b + 1] = q3; }
    else if (r == 3) { list[b] = q1; list[b + 1] = q2; list[b + 2] = q3; }
  }
};

// ------------------------------------------------------------------------------------------------
// One member.  Returns true when its token stream is complete, false when the member must be
// redone by the sequential kernel.
// ------------------------------------------------------------------------------------------------
__device__ inline bool decode_member(const DMember &mem, int fmt, P1Rec &rec, Smem &sm,
                                     uint32_t *__restrict__ slabs, uint32_t nslabs, uint32_t *slab_counter, int tid) {
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
  if (tid == 0) sm.fail = 0;
  unsigned long long A = 0;    // output bytes so far
  uint32_t first_slab = NO_SLAB, prev_slab = NO_SLAB;
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
          for (int e = lane; e < 128; e += 32) sm.lut_cl[e] = canon_lookup(sm.c_cl, sm.sorted_cl, (uint32_t)e, 1, 7);
          __syncwarp();
          if (lane == 0) {
            // the code lengths themselves: one lane, table driven (deflate.lisp:626-669)
            uint32_t p = pos + 14 + 3 * ncl;
            int idx = 0, lastlen = 0xff;
            const int total = hlit + hdist;
            Bits hb;
            bits_init(hb, in, p);
            while (idx < total) {
              bits_refill(hb, in);
              uint32_t w = (uint32_t)hb.bb;
              uint32_t r = sm.lut_cl[w & 127];
              if (!r) { err = TBZ_ERR_INVALID_SYMBOL; break; }
              int L = r & 15, sym = r >> 4;
              int xb = sym < 16 ? 0 : sym == 16 ? 2 : sym == 17 ? 3 : 7;
              if (p + L + xb > in.end) { err = TBZ_INPUT_UNDERRUN; break; }
              uint32_t extra = (w >> L) & ((1u << xb) - 1);
              p += L + xb;
              bits_skip(hb, L + xb);
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
      // ---- a slab for this round's token lists
      if (tid == 0) {
        uint32_t s = atomicAdd(slab_counter, 1u);
        if (s >= nslabs) { s = NO_SLAB; sm.fail = 1; }
        sm.slab = s;
      }
      // ---- geometry of this round
      const uint32_t P0 = pos;
      const uint32_t winbase = P0 & ~31u;
      uint32_t S = ((in.end - winbase + NT - 1) / NT + 31) & ~31u;
      if (S > S_MAX) S = S_MAX;
      if (S < S_MIN) S = S_MIN;
      const uint32_t winend = winbase + S * NT;
      const uint32_t bmwords = (S * NT) >> 5;
      for (uint32_t w = tid; w < bmwords; w += NT) sm.bitmap[w] = 0;
      __syncthreads();
      if (sm.fail) return false;
      uint32_t *slab = slabs + (size_t)sm.slab * SLAB_WORDS;
      SlabHdr *sh = reinterpret_cast<SlabHdr *>(slab);
      TokW tw;
      tw.list = slab + sizeof(SlabHdr) / 4 + tid * TOKCAP;
      tw.q0 = tw.q1 = tw.q2 = tw.q3 = 0;

      // ---- 1a: speculative decode of the lane's sub-chunk
      const uint32_t cstart = winbase + S * tid, cend = cstart + S;
      uint32_t p = tid == 0 ? P0 : cstart;
      uint32_t k = 0, ob = 0, nx = NT;
      int st = ST_IDLE;
      Bits b;
      if (p < in.end) {
        bits_init(b, in, p);
        uint32_t curw = (p - winbase) >> 5, acc = 0;
        for (;;) {
          const uint32_t rel = p - winbase;
          if (p >= cend) {
            // Past the own sub-chunk: look for a bit a later lane marked.  That lane is usually far
            // ahead in its own sub-chunk by now; a mark that is not visible yet only delays the
            // match (any later common token start is as good), and after SYNC_SLACK bits the lane
            // parks until the barrier below has made every mark visible.
            if (curw != NO_SLAB) { sm.bitmap[curw] = acc; curw = NO_SLAB; }     // own marks are complete
            if (p >= winend || p >= cend + SYNC_SLACK) { st = ST_END; break; }
            if ((*(volatile uint32_t *)&sm.bitmap[rel >> 5] >> (rel & 31)) & 1u) { st = ST_SYNC; nx = rel / S; break; }
          }
          if (k >= TOKCAP) { st = ST_CAP; break; }
          if (p < cend) {
            const uint32_t wi = rel >> 5;
            if (wi != curw) { sm.bitmap[curw] = acc; acc = 0; curw = wi; }
            acc |= 1u << (rel & 31);
          }
          uint32_t tok, nb, ol;
          const int kind = decode_token(b, in, sm, tok, nb, ol);
          if (kind == 3 || p + nb > in.end) { st = ST_BAD; break; }
          tw.push(tok, k);
          k++; p += nb; ob += ol;
          if (kind == 2) { st = ST_EOB; break; }
        }
        if (curw != NO_SLAB) sm.bitmap[curw] = acc;
      }
      __syncthreads();
      // ---- 1b: lanes still looking for their synchronisation point go on with every mark visible
      if (st == ST_END) {
        for (;;) {
          if (p >= winend) break;                                 // round ends here, block continues
          const uint32_t rel = p - winbase;
          if ((sm.bitmap[rel >> 5] >> (rel & 31)) & 1u) { st = ST_SYNC; nx = rel / S; break; }
          if (k >= TOKCAP) { st = ST_CAP; break; }
          uint32_t tok, nb, ol;
          const int kind = decode_token(b, in, sm, tok, nb, ol);
          if (kind == 3 || p + nb > in.end) { st = ST_BAD; break; }
          tw.push(tok, k);
          k++; p += nb; ob += ol;
          if (kind == 2) { st = ST_EOB; break; }
        }
      }
      tw.finish(k);
      sm.nxt[0][tid] = (uint16_t)nx;
      sm.truth[tid] = tid == 0;
      if (tid == 0) { sm.nxt[0][NT] = NT; sm.truth[NT] = 0; }
      __syncthreads();
      // ---- 1c: lanes reachable from lane 0 through "synchronised into" edges are proven
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
        else { sm.term_status = st; sm.term_pos = p; }
      }
      if (tid == 0) sm.entry[0] = P0;
      __syncthreads();
      if (sm.term_status == ST_BAD || sm.term_status == ST_IDLE) return false;
      // ---- tokens decoded before the entry point are dropped: count and size them
      uint32_t g = 0, gb = 0;
      if (proven && tid != 0) {
        const uint32_t r0 = cstart - winbase, r1 = sm.entry[tid] - winbase;   // r1 in [r0, r0 + S)
        for (uint32_t w = r0 >> 5; w <= (r1 >> 5); w++) {
          uint32_t bits = sm.bitmap[w];
          if (w == (r1 >> 5)) bits &= (1u << (r1 & 31)) - 1u;
          g += __popc(bits);
        }
        for (uint32_t i = 0; i < g; i++) gb += tok_outlen(tw.list[i]);
      }
      sh->gn[tid] = proven ? (g | (k << 16)) : 0u;
      uint32_t total, ntok;
      {
        // exclusive scan of the proven token counts (flat token order for phase two) + output total
        const uint32_t cnt = proven ? k - g : 0u;
        uint32_t x = cnt, y = proven ? ob - gb : 0u;
#pragma unroll
        for (int sft = 1; sft < 32; sft <<= 1) {
          const uint32_t u = __shfl_up_sync(TBZ_FULL, x, sft);
          if (lane >= sft) x += u;
        }
#pragma unroll
        for (int sft = 16; sft; sft >>= 1) y += __shfl_xor_sync(TBZ_FULL, y, sft);
        if (lane == 31) { sm.wscan[warp] = x; sm.wscan2[warp] = y; }
        __syncthreads();
        uint32_t off = 0;
        ntok = 0; total = 0;
#pragma unroll
        for (int w = 0; w < NWARP; w++) { const uint32_t c = sm.wscan[w]; if (w < warp) off += c; ntok += c; total += sm.wscan2[w]; }
        sh->tb[tid] = off + x - cnt;
        __syncthreads();
      }
      A += total;
      if (A > mem.out_cap || A >= (1ull << 32)) return false;       // overflow: sequential kernel
      if (tid == 0) {
        sh->next = NO_SLAB; sh->out_bytes = total; sh->ntokens = ntok;
        if (prev_slab != NO_SLAB) reinterpret_cast<SlabHdr *>(slabs + (size_t)prev_slab * SLAB_WORDS)->next = sm.slab;
      }
      if (first_slab == NO_SLAB) first_slab = sm.slab;
      prev_slab = sm.slab;
      // ---- how did the round end?
      pos = sm.term_pos;
      if (sm.term_status == ST_EOB) block_done = true;
      __syncthreads();
    }
  }
  if (tid == 0) {
    rec.first_slab = first_slab;
    rec.out_len = (uint32_t)A;
    rec.end_pos = pos;
    rec.status = 1;
  }
  return true;
}

}  // namespace tbzfast
