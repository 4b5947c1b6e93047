// inflate_decode.cuh — phase one of the batched fast path: Huffman decode into a token stream.
// One WARP per member, every lane a decode lane; a CTA holds WPC independent warps.
//
// Inside one Huffman block the compressed bits are cut into 32 equal sub-chunks of S bits.  Lane i
// starts decoding at the first bit of sub-chunk i *speculatively* (only lane 0 is known to start
// on a symbol boundary) and relies on the self-synchronisation of Huffman streams:
//   1a  every lane decodes its sub-chunk into its own token list (global memory, 16-byte stores)
//       and records a checkpoint (bit position, output bytes so far) every CKSTEP tokens in
//       shared memory
//   1b  every lane keeps decoding past its sub-chunk end until the start of one of its tokens
//       coincides with a checkpoint of a later lane — from there on both decodes are identical
//       (same tables, same bit), so the rest of that lane's list is proven correct and the
//       tokens that lane decoded before the checkpoint are dropped
//   1c  a walk over "who synchronised into whom" from lane 0 gives the proven lanes, their first
//       proven token and the round's output size
// Measured on the level-6 text of BASELINE config 2: a mis-aligned start re-synchronises after
// 117 bits on average (p99 595), against S ~ 3000 - 4000 bits per lane, so ~95 % of decode work is kept.
// Phase two (inflate_copy.cuh) turns the token stream into bytes.  Stored blocks travel as literal
// tokens.  Anything this kernel cannot prove clean (malformed codes, truncated input, too small an output buffer ...)
// is queued for the sequential kernel, which reproduces the reference's exact verdict.
// Replaces deflate.lisp:465-509,673-702 (decode) and huffman-tree.lisp:99-218 (tables).
#pragma once
#include <cstddef>
#include "tbz_device.cuh"

namespace tbzfast {

#ifndef TBZ_DEC_TOKCAP
#define TBZ_DEC_TOKCAP 736
#endif
#ifndef TBZ_DEC_CKSTEP
#define TBZ_DEC_CKSTEP 32
#endif
#ifndef TBZ_DEC_SMAX
#define TBZ_DEC_SMAX 4000
#endif
constexpr int WPC = 4;                  // warps (members in flight) per CTA
constexpr int NT = WPC * 32;
constexpr int NL = 32;                  // decode lanes per member
constexpr int KLL = 10, KD = 9;         // root table bits: lit/len, distance
constexpr int TOKCAP = TBZ_DEC_TOKCAP;            // tokens a lane may emit per round (sub-chunk + overrun)
constexpr int CKSTEP = TBZ_DEC_CKSTEP;              // a checkpoint every CKSTEP tokens
constexpr int NCK = TOKCAP / CKSTEP;
constexpr uint32_t S_MAX = TBZ_DEC_SMAX, S_MIN = 64;   // sub-chunk size in bits (13-bit field in a checkpoint)
constexpr uint32_t CK_NONE = 0xffffffffu;

// A slab holds the token lists of one round: header, then NL lists of TOKCAP tokens.
struct SlabHdr {
  uint32_t next;        // next slab of the member, or NO_SLAB
  uint32_t out_bytes;   // output bytes of the round
  uint32_t pad[2];
  uint32_t fc[NL];      // per lane: first proven token | (number of proven tokens << 16); 0 = nothing
};
constexpr uint32_t SLAB_HDR_WORDS = sizeof(SlabHdr) / 4;
constexpr uint32_t SLAB_WORDS = SLAB_HDR_WORDS + NL * TOKCAP;
constexpr uint32_t NO_SLAB = 0xffffffffu;

// what phase one leaves per member for phase two
struct P1Rec {
  uint32_t status;      // 1 = token stream complete, 0 = member goes to the sequential kernel
  uint32_t first_slab;
  uint32_t out_len;     // total output bytes
  uint32_t end_pos;     // bit position (relative to the 4-byte aligned input base) after the last block
};

constexpr uint32_t E_LONG = 0xc000u, E_INVALID = 0x8000u;   // table specials (code length 0)
// token: one literal = byte value; two literals = TOK_LIT2 | second << 8 | first;
// match = TOK_MATCH | (distance - 1) << 8 | (length - 3).  End of block is not a token.
constexpr uint32_t TOK_MATCH = 0x80000000u, TOK_LIT2 = 0x40000000u;

enum { ST_IDLE = 0, ST_OVER, ST_END, ST_SYNC, ST_EOB, ST_CAP, ST_BAD };

struct Canon16 { uint16_t first[16], count[16], base[16]; uint16_t maxlen, nsyms; };

struct HdrScratch {                      // only alive while a block header is parsed
  uint32_t lut_cl[128];
  Canon16 c_cl;
  uint16_t sorted_cl[32];
  uint8_t lens[352];                     // [0,19) code-length code, [32,352) lit/len + distance
  uint16_t run[16];
};
struct WSmem {                           // one per warp
  uint16_t lut_ll[1 << KLL];             // 16-bit entries: 7 KB per warp in all, 8 CTAs (32 warps) per SM
  uint16_t lut_d[1 << KD];               // (directly behind lut_ll: the second lookup of a token indexes either)
  union {
    uint32_t ckpt[NCK][NL];              // [checkpoint][lane]: (bit offset in the sub-chunk) | (output bytes << 13)
    HdrScratch h;
  };
  Canon16 c_ll, c_d;
  uint16_t sorted_ll[288], sorted_d[32];
  uint16_t lenbase[32], distbase[32];    // constants.lisp:36-61 base values, next to the tables that index them
#ifdef TBZ_DEC_PAD
  uint8_t pad[TBZ_DEC_PAD];            // occupancy experiments
#endif
};
static_assert(sizeof(HdrScratch) <= sizeof(uint32_t) * NCK * NL, "header scratch must fit under the checkpoints");
static_assert(offsetof(WSmem, lut_d) == offsetof(WSmem, lut_ll) + sizeof(uint16_t) * (1 << KLL), "lut_d must follow lut_ll");
static_assert(sizeof(WSmem) * WPC + 1024 <= 233472 / 8, "eight CTAs per SM");

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

// ---- table entries (16 bit) -----------------------------------------------------------------
// lit/len: [3:0] code length (0 = special), [6:4] extra bits, then
//            literal      : bit 15 = 0, [14:7] the byte
//            length       : bit 15 = 1, bit 14 = 0, [11:7] length symbol - 257 (base value in lenbase[])
//            end of block : bits 15 and 14 = 1
// dist   : [3:0] code length, [7:4] extra bits, [12:8] distance symbol (base value in distbase[])
__device__ __forceinline__ uint32_t ll_entry(uint32_t sym, uint32_t L) {
  if (sym < 256) return (sym << 7) | L;
  if (sym == 256) return 0xc000u | L;
  if (sym > 285) return E_INVALID;                          // huffman-tree.lisp:176-177
  return 0x8000u | ((sym - 257) << 7) | ((uint32_t)c_len_extra[sym - 257] << 4) | L;
}
__device__ __forceinline__ uint32_t d_entry(uint32_t sym, uint32_t L) {
  if (sym > 29) return E_INVALID;                           // huffman-tree.lisp:172-175
  return (sym << 8) | ((uint32_t)c_dist_extra[sym] << 4) | L;
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
#ifndef TBZ_EMU
    if ((b.wn & 31u) == 0u && b.wn + 32u < in.nwords) asm volatile("prefetch.global.L1 [%0];" ::"l"(in.w + b.wn + 32));
#endif
  }
}
__device__ __forceinline__ void bits_skip(Bits &b, uint32_t n) { b.bb >>= n; b.bc -= n; }

// One token.  Two table lookups on one instruction path for every lane: the first decodes a
// lit/len symbol; the second decodes the distance after a length, or — after a literal — a second
// literal, which is consumed only if it is one (two literals travel in one token).  The lanes of a
// warp do not diverge on the token kind.
// Returns 0 literal(s), 1 match, 2 end of block, 3 invalid code.
__device__ __forceinline__ int decode_token(Bits &b, const In &in, const WSmem &sm, uint32_t &tok, uint32_t &nbits, uint32_t &olen) {
  nbits = 0;
  bits_refill(b, in);                                        // >= 33 bits
  uint32_t w = (uint32_t)b.bb;
  uint32_t e = sm.lut_ll[w & ((1u << KLL) - 1)];
  if ((e & 15) == 0) {
    if (e != E_LONG) return 3;
        uint32_t r = canon_lookup(sm.c_ll, sm.sorted_ll, w, KLL + 1, 15);
    if (!r) return 3;
    e = ll_entry(r >> 4, r & 15);
    if ((e & 15) == 0) return 3;
  }
  const uint32_t L = e & 15, xb = (e >> 4) & 7, t2 = e >> 14;
  const uint32_t kind = (t2 < 1u ? 1u : t2) - 1u;            // 0 literal, 1 length, 2 end of block
  const bool ism = kind == 1;
  const uint32_t pay = (e >> 7) & 255u;                      // the byte, or the length symbol
  const uint32_t lb = sm.lenbase[pay & 31u];                  // (loaded by every lane: no divergence)
  const uint32_t val = ism ? lb + ((w >> L) & ((1u << xb) - 1)) : pay;   // match length / the byte
  const uint32_t n1 = L + xb;
  bits_skip(b, n1);
  if (ism && b.bc < 28) bits_refill(b, in);                  // a literal leaves >= 18 bits: enough for any code
  const uint32_t w2 = (uint32_t)b.bb;
  // one load for either table: lut_d lies directly behind lut_ll
  const uint16_t *const lut = sm.lut_ll;
  uint32_t d = lut[ism ? (1u << KLL) + (w2 & ((1u << KD) - 1)) : (w2 & ((1u << KLL) - 1))];
  if (ism && (d & 15) == 0) {
    if (d != E_LONG) return 3;
        uint32_t r = canon_lookup(sm.c_d, sm.sorted_d, w2, KD + 1, 15);
    if (!r) return 3;
    d = d_entry(r >> 4, r & 15);
    if ((d & 15) == 0) return 3;
  }
  // after a literal: keep the second symbol only if it is a literal with a short code
  const bool two = kind == 0 && (d & 15) != 0 && (d >> 15) == 0;
  const uint32_t DL = (ism || two) ? (d & 15) : 0u;
  const uint32_t dxb = ism ? (d >> 4) & 15 : 0u;             // a literal has no extra bits
  const uint32_t db = sm.distbase[(d >> 8) & 31u];
  const uint32_t dv = ism ? db + ((w2 >> DL) & ((1u << dxb) - 1)) : (d >> 7) & 255u;   // distance / second literal
  const uint32_t n2 = DL + dxb;
  bits_skip(b, n2);
  tok = ism ? (TOK_MATCH | ((dv - 1) << 8) | (val - 3)) : two ? (TOK_LIT2 | (dv << 8) | val) : val;
  nbits = n1 + n2;
  olen = ism ? val : (kind == 0 ? 1u + (two ? 1u : 0u) : 0u);
  return (int)kind;
}

__device__ __forceinline__ uint32_t tok_outlen(uint32_t t) { return (t & TOK_MATCH) ? (t & 255u) + 3u : 1u + ((t >> 30) & 1u); }

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

// Candidate block starts of a split member, ascending absolute bit positions (chunk decode only):
// the decode of a chunk ends on the first candidate it lands on EXACTLY; candidates it passes over
// were not block starts (false positives of the search) and are simply skipped.
struct StopList {
  const unsigned long long *starts; uint32_t n, next; unsigned long long base;   // base: absolute bit of in.w
  __device__ __forceinline__ uint32_t rel(uint32_t i) const {
    if (i >= n) return 0xffffffffu;
    const unsigned long long r = starts[i] - base;
    return r > 0xf0000000ull ? 0xf0000000u : (uint32_t)r;
  }
};

// ------------------------------------------------------------------------------------------------
// Blocks from bit `pos` (a block start) on, one warp.  Stops after the final block, or — for a
// chunk of a split member — after the first block that ends at or beyond `stop_bit`.  Returns
// true when the token stream is complete (rec filled in), false when the caller must fall back.
// Every return value is warp-uniform.
// ------------------------------------------------------------------------------------------------
__device__ inline bool decode_blocks(const In &in, uint32_t pos, uint32_t stop_bit, unsigned long long out_cap, P1Rec &rec,
                                     WSmem &sm, uint32_t *__restrict__ slabs, uint32_t nslabs, uint32_t *slab_counter, int lane,
                                     const StopList *stops = nullptr) {
  uint32_t stop_i = stops ? stops->next : 0u;
  unsigned long long A = 0;    // output bytes so far
  sm.lenbase[lane] = c_len_base[lane]; sm.distbase[lane] = c_dist_base[lane];
  uint32_t first_slab = NO_SLAB, prev_slab = NO_SLAB;
  uint32_t prev_block_bits = 0;   // size of the previous block of this member: predicts this one
  bool last = false;

  while (!last) {
    // ================= block header (deflate.lisp:518-528, :577-669) =================
    if (in.end - pos < 3) return false;
    const uint32_t hdr = peek32(in, pos) & 7;
    pos += 3;
    last = hdr & 1;
    const uint32_t btype = hdr >> 1;
    int hlit, hdist;
    __syncwarp();
    if (btype == 1) {
      hlit = 288; hdist = 32;
      for (int i = lane; i < 320; i += 32) sm.h.lens[32 + i] = i < 144 ? 8 : i < 256 ? 9 : i < 280 ? 7 : i < 288 ? 8 : 5;
    } else if (btype == 2) {
      if (in.end - pos < 14) return false;
      const uint32_t v = peek32(in, pos);
      hlit = (v & 31) + 257; hdist = ((v >> 5) & 31) + 1;
      const int ncl = ((v >> 10) & 15) + 4;
      if (in.end - pos < 14u + 3u * ncl) return false;
      if (lane < 19) sm.h.lens[lane] = 0;
      __syncwarp();
      if (lane < ncl) sm.h.lens[c_clen_order[lane]] = peek32(in, pos + 14 + 3 * lane) & 7;
      __syncwarp();
      int err = warp_canon(sm.h.lens, 19, sm.h.c_cl, sm.h.sorted_cl, sm.h.run, lane);
      if (!err && sm.h.c_cl.nsyms == 0) err = TBZ_ERR_INVALID_SYMBOL;
      if (err) return false;
      // entry: [3:0] code length, [7:4] extra bits, [12:8] symbol; 0 = no code
      for (int e = lane; e < 128; e += 32) {
        const uint32_t r = canon_lookup(sm.h.c_cl, sm.h.sorted_cl, (uint32_t)e, 1, 7);
        const uint32_t sym = r >> 4;
        const uint32_t xb = sym < 16 ? 0 : sym == 16 ? 2 : sym == 17 ? 3 : 7;
        sm.h.lut_cl[e] = r ? ((r & 15) | (xb << 4) | (sym << 8)) : 0;
      }
      __syncwarp();
      uint32_t p = pos + 14 + 3 * ncl;
      if (lane == 0) {
        // the code lengths themselves: one lane, table driven (deflate.lisp:626-669)
        int idx = 0, lastlen = 0xff;
        const int total = hlit + hdist;
        Bits hb;
        bits_init(hb, in, p);
        while (idx < total) {
          bits_refill(hb, in);
          const uint32_t w = (uint32_t)hb.bb;
          const uint32_t r = sm.h.lut_cl[w & 127];
          if (!r) { err = 1; break; }
          const uint32_t L = r & 15, xb = (r >> 4) & 15, sym = r >> 8;
          if (p + L + xb > in.end) { err = 1; break; }
          p += L + xb;
          if (sym < 16) {
            sm.h.lens[32 + idx] = (uint8_t)sym; idx++; lastlen = (int)sym;
            bits_skip(hb, L);
            continue;
          }
          const uint32_t extra = (w >> L) & ((1u << xb) - 1);
          bits_skip(hb, L + xb);
          int rep, val;
          if (sym == 16) { if (lastlen >= 16) { err = 1; break; } rep = 3 + extra; val = lastlen; }
          else { rep = (sym == 17 ? 3 : 11) + extra; val = 0; lastlen = 0; }
          if (idx + rep > total) { err = 1; break; }
          for (int k = 0; k < rep; k++) sm.h.lens[32 + idx + k] = (uint8_t)val;
          idx += rep;
        }
      }
      err = __shfl_sync(TBZ_FULL, err, 0);
      if (err) return false;
      pos = __shfl_sync(TBZ_FULL, p, 0);
    } else if (btype == 0) {
      // ================= stored block (deflate.lisp:532-573) =================
      // LEN, NLEN, then LEN bytes as they are.  They travel as literal tokens, two bytes each, the
      // lanes taking consecutive slices; anything irregular is left to the sequential kernel.
      pos = (pos + 7u) & ~7u;
      if (in.end < pos || in.end - pos < 32u) return false;
      const uint32_t v = peek32(in, pos);
      uint32_t slen = v & 0xffffu;
      if ((slen ^ 0xffffu) != (v >> 16)) return false;                 // deflate.lisp:535
      pos += 32u;
      if (in.end - pos < 8u * slen) return false;                      // the input ends inside the block
      uint32_t bp = pos >> 3;                                          // byte offset of the payload from in.w
      pos += 8u * slen;
      while (slen) {
        uint32_t slab_id = 0;
        if (lane == 0) slab_id = atomicAdd(slab_counter, 1u);
        slab_id = __shfl_sync(TBZ_FULL, slab_id, 0);
        if (slab_id >= nslabs) return false;
        uint32_t *slab = slabs + (size_t)slab_id * SLAB_WORDS;
        SlabHdr *sh = reinterpret_cast<SlabHdr *>(slab);
        const uint32_t nb = slen < (uint32_t)(NL * TOKCAP * 2) ? slen : (uint32_t)(NL * TOKCAP * 2);
        const uint32_t per = ((nb + NL - 1) / NL + 1u) & ~1u;          // bytes per lane, even
        const uint32_t lo = per * lane < nb ? per * lane : nb, hi = lo + per < nb ? lo + per : nb;
        const uint32_t cnt = (hi - lo + 1u) / 2u;
        uint32_t *list = slab + SLAB_HDR_WORDS + lane * TOKCAP;
        for (uint32_t t = 0; t < cnt; t++) {
          const uint32_t a = bp + lo + 2u * t;
          const uint32_t b0 = byte_at(in, a);
          list[t] = lo + 2u * t + 1u < hi ? (TOK_LIT2 | (byte_at(in, a + 1u) << 8) | b0) : b0;
        }
        A += nb;
        if (A > out_cap || A >= (1ull << 32)) return false;            // overflow: sequential kernel
        sh->fc[lane] = cnt << 16;
        if (lane == 0) {
          sh->next = NO_SLAB; sh->out_bytes = nb;
          if (prev_slab != NO_SLAB) reinterpret_cast<SlabHdr *>(slabs + (size_t)prev_slab * SLAB_WORDS)->next = slab_id;
        }
        if (first_slab == NO_SLAB) first_slab = slab_id;
        prev_slab = slab_id;
        bp += nb; slen -= nb;
        __syncwarp();
      }
      prev_block_bits = 0;
      if (stops) {
        while (pos > stop_bit) stop_bit = stops->rel(++stop_i);
        if (pos == stop_bit) break;
      } else if (pos >= stop_bit) break;
      continue;
    } else {
      return false;                        // reserved block type: sequential kernel
    }
    __syncwarp();
    // ================= tables (huffman-tree.lisp:99-218) =================
    if (warp_canon(sm.h.lens + 32, hlit, sm.c_ll, sm.sorted_ll, sm.h.run, lane)) return false;
    if (warp_canon(sm.h.lens + 32 + hlit, hdist, sm.c_d, sm.sorted_d, sm.h.run, lane)) return false;
    if (sm.c_ll.nsyms == 0) return false;
    for (int e = lane; e < (1 << KLL); e += 32) {
      uint32_t r = canon_lookup(sm.c_ll, sm.sorted_ll, (uint32_t)e, 1, KLL);
      sm.lut_ll[e] = (uint16_t)(r ? ll_entry(r >> 4, r & 15) : (sm.c_ll.maxlen > KLL ? E_LONG : E_INVALID));
    }
    for (int e = lane; e < (1 << KD); e += 32) {
      uint32_t r = canon_lookup(sm.c_d, sm.sorted_d, (uint32_t)e, 1, KD);
      sm.lut_d[e] = (uint16_t)(r ? d_entry(r >> 4, r & 15) : (sm.c_d.maxlen > KD ? E_LONG : E_INVALID));
    }
    __syncwarp();

    // ================= rounds over the block's compressed bits =================
    // The end of the block is unknown: assume it is about as long as the previous one (libz cuts
    // blocks by symbol count), else that it runs to the end of the input.
    const uint32_t data_start = pos;
    uint32_t expect = in.end - pos;
    if (prev_block_bits && prev_block_bits + prev_block_bits / 16 < expect) expect = prev_block_bits + prev_block_bits / 16;
    bool block_done = false;
    uint32_t shrink = 0;
    while (!block_done) {
      // ---- a slab for this round's token lists
      uint32_t slab_id = 0;
      if (lane == 0) slab_id = atomicAdd(slab_counter, 1u);
      slab_id = __shfl_sync(TBZ_FULL, slab_id, 0);
      if (slab_id >= nslabs) return false;
      // ---- geometry of this round
      const uint32_t P0 = pos;
      uint32_t left = in.end - P0;
      if (expect > pos - data_start && expect - (pos - data_start) < left) left = expect - (pos - data_start);
      const uint32_t nrounds = (left + NL * S_MAX - 1) / (NL * S_MAX);
      uint32_t S = ((left + nrounds - 1) / nrounds + NL - 1) / NL;
      S >>= shrink;
      if (S > S_MAX) S = S_MAX;
      if (S < S_MIN) S = S_MIN;
      const uint32_t winend = P0 + S * NL;
      for (int c = 0; c < NCK; c++) sm.ckpt[c][lane] = CK_NONE;
      uint32_t *slab = slabs + (size_t)slab_id * SLAB_WORDS;
      SlabHdr *sh = reinterpret_cast<SlabHdr *>(slab);
      TokW tw;
      tw.list = slab + SLAB_HDR_WORDS + lane * TOKCAP;
      tw.q0 = tw.q1 = tw.q2 = tw.q3 = 0;
      __syncwarp();

      // ---- 1a: speculative decode of the lane's sub-chunk.  The loop is kept warp-converged (one
      // vote per iteration, the body predicated on `act`): lanes that leave a loop through a break
      // are not guaranteed to re-converge, and a split warp issues the whole body once per fragment.
      const uint32_t cstart = P0 + S * lane, cend = cstart + S;
      uint32_t p = cstart;
      uint32_t k = 0, ob = 0;
      int st = ST_IDLE;
      Bits b;
      bool act = p < in.end;
      if (act) bits_init(b, in, p);
      while (__any_sync(TBZ_FULL, act)) {
        if (act) {
          if (p >= cend) { st = ST_OVER; act = false; }
          else {
            if ((k & (CKSTEP - 1)) == 0) {
              if (k >= TOKCAP) { st = ST_CAP; act = false; }
              else sm.ckpt[k / CKSTEP][lane] = (p - cstart) | (ob << 13);
            }
            if (act) {
              uint32_t tok, nb, ol;
              const int kind = decode_token(b, in, sm, tok, nb, ol);
              p += nb;
              if (kind >= 2) { st = kind == 2 ? ST_EOB : ST_BAD; act = false; }
              else { tw.push(tok, k); k++; ob += ol; }
            }
          }
        }
      }
      __syncwarp();
      // ---- 1b: past the own sub-chunk: decode on until a token start coincides with a checkpoint
      // of the lane whose sub-chunk the position lies in.  Same converged loop; an iteration either
      // decodes one token or looks up the next place a synchronisation can happen (`tgt`).
      uint32_t nx = 0, g_sync = 0, ob_sync = 0;
      {
        uint32_t j = lane, jend = cend, c = 0, tgt = 0;
        act = st == ST_OVER;
        while (__any_sync(TBZ_FULL, act)) {
          if (act) {
            if (p >= tgt) {
              if (p >= winend) { st = ST_END; act = false; }
              else {
                while (p >= jend) { j++; jend += S; c = 0; }
                const uint32_t rel = p - (jend - S);
                uint32_t ck = CK_NONE;
                while (c < NCK && ((ck = sm.ckpt[c][j]) & 0x1fffu) < rel) c++;
                if (c < NCK && ck != CK_NONE && (ck & 0x1fffu) == rel) {
                  st = ST_SYNC; act = false; nx = j; g_sync = c * CKSTEP; ob_sync = ck >> 13;
                } else {
                  tgt = jend;
                  if (c < NCK && ck != CK_NONE) tgt = (jend - S) + (ck & 0x1fffu);
                }
              }
            } else if (k >= TOKCAP) { st = ST_CAP; act = false; }
            else {
              uint32_t tok, nb, ol;
              const int kind = decode_token(b, in, sm, tok, nb, ol);
              p += nb;
              if (kind >= 2) { st = kind == 2 ? ST_EOB : ST_BAD; act = false; }
              else { tw.push(tok, k); k++; ob += ol; }
            }
          }
        }
      }
      // a lane that decoded past the end of the input has nothing proven to offer (zeros are read there)
      if (st != ST_IDLE && p > in.end) st = ST_BAD;
      tw.finish(k);
      __syncwarp();
      // ---- 1c: lanes reachable from lane 0 through "synchronised into" edges are proven
      uint32_t my_g = 0, my_gob = 0;
      bool proven = false;
      int term_st;
      uint32_t term_pos;
      {
        int cur = 0;
        for (;;) {
          if (lane == cur) proven = true;
          const int st_c = __shfl_sync(TBZ_FULL, st, cur);
          if (st_c != ST_SYNC) { term_st = st_c; term_pos = __shfl_sync(TBZ_FULL, p, cur); break; }
          const uint32_t nx_c = __shfl_sync(TBZ_FULL, nx, cur);
          const uint32_t g_c = __shfl_sync(TBZ_FULL, g_sync, cur), gob_c = __shfl_sync(TBZ_FULL, ob_sync, cur);
          if ((uint32_t)lane == nx_c) { my_g = g_c; my_gob = gob_c; }
          cur = (int)nx_c;
        }
      }
      if (term_st == ST_BAD || term_st == ST_IDLE || term_st == ST_OVER) return false;
      // a lane that ran into its token cap ends the round early: use shorter sub-chunks from here on
      if (term_st == ST_CAP && shrink < 6) shrink++;
      const uint32_t cnt = proven ? k - my_g : 0u;
      uint32_t total = proven ? ob - my_gob : 0u;
#pragma unroll
      for (int sft = 16; sft; sft >>= 1) total += __shfl_xor_sync(TBZ_FULL, total, sft);
      A += total;
      if (A > out_cap || A >= (1ull << 32)) return false;            // overflow: sequential kernel
      sh->fc[lane] = cnt ? (my_g | (cnt << 16)) : 0u;
      if (lane == 0) {
        sh->next = NO_SLAB; sh->out_bytes = total;
        if (prev_slab != NO_SLAB) reinterpret_cast<SlabHdr *>(slabs + (size_t)prev_slab * SLAB_WORDS)->next = slab_id;
      }
      if (first_slab == NO_SLAB) first_slab = slab_id;
      prev_slab = slab_id;
      // ---- how did the round end?
      if (term_pos <= pos && term_st != ST_EOB) return false;     // no progress (cannot happen; guards the loop)
      pos = term_pos;
      if (term_st == ST_EOB) block_done = true;
      __syncwarp();
    }
    prev_block_bits = pos - data_start;
    if (stops) {                              // skip the candidates this block ran past; stop on an exact landing
      while (pos > stop_bit) stop_bit = stops->rel(++stop_i);
      if (pos == stop_bit) break;
    } else if (pos >= stop_bit) break;
  }
  if (lane == 0) {
    rec.first_slab = first_slab;
    rec.out_len = (uint32_t)A;
    rec.end_pos = pos;
    rec.status = last ? 1u : 2u;           // 2: stopped at a block boundary before the final block
  }
  return true;
}

// Where a member's deflate body starts: the input window (4-byte aligned base at or below the first byte) and the
// wrapper header (zlib.lisp:108-126, gzip.lisp:113-240).  false: the sequential kernel owns the verdict.
__device__ inline bool member_start(const DMember &mem, int fmt, In &in, uint32_t &pos) {
  {
    uintptr_t a = (uintptr_t)mem.in;
    uint32_t mis = (uint32_t)(a & 3);
    in.w = (const uint32_t *)(a - mis);
    in.pos0 = mis * 8;
    if (mem.in_len >= (1ull << 28)) return false;
    in.end = (mis + (uint32_t)mem.in_len) * 8;
    in.nwords = (in.end + 31) >> 5;
  }
  pos = in.pos0;
  if (fmt == TBZ_ZLIB) {
    if (in.end - pos < 16) return false;
    uint32_t cmf = byte_at(in, pos >> 3), flg = byte_at(in, (pos >> 3) + 1);
    if ((cmf * 256 + flg) % 31 || (cmf & 15) != 8 || (cmf >> 4) > 7 || (flg & 32)) return false;
    pos += 16;
  } else if (fmt == TBZ_GZIP) {
    if (in.end - pos < 80) return false;
    uint32_t bp = pos >> 3;
    if (byte_at(in, bp) != 0x1f || byte_at(in, bp + 1) != 0x8b || byte_at(in, bp + 2) != 8) return false;
    const uint32_t flg = byte_at(in, bp + 3);
    if (flg & 0xe2u) return false;                      // reserved bits (an error) or a header CRC: sequential kernel
    // optional fields (gzip.lisp:178-240): FEXTRA, FNAME, FCOMMENT are skipped here — `gzip file` always
    // writes a name; every lane walks the same bytes
    const uint32_t endb = in.end >> 3;
    uint32_t q = bp + 10;
    if (flg & 4u) {
      if (q + 2 > endb) return false;
      q += 2u + (byte_at(in, q) | (byte_at(in, q + 1) << 8));
    }
    for (uint32_t field = 8u; field <= 16u; field <<= 1) {
      if (flg & field) {
        while (q < endb && byte_at(in, q)) q++;
        q++;
      }
    }
    if (q >= endb) return false;                        // the header does not end inside the input
    pos = q * 8;
  }
  return true;
}

// One member, one warp: wrapper header, then every block.
__device__ inline bool decode_member(const DMember &mem, int fmt, P1Rec &rec, WSmem &sm,
                                     uint32_t *__restrict__ slabs, uint32_t nslabs, uint32_t *slab_counter, int lane) {
  In in;
  uint32_t pos;
  if (!member_start(mem, fmt, in, pos)) return false;
  return decode_blocks(in, pos, 0xffffffffu, mem.out_cap, rec, sm, slabs, nslabs, slab_counter, lane);
}

// ------------------------------------------------------------------------------------------------
// Split decode of one large member (pugz-style): where does the first dynamic block at or after
// bit `from` start?  32 candidate bit positions per step, one per lane: block type, HLIT / HDIST
// range and completeness of the code-length code are tested by every lane for its own candidate.
// The survivors (2 - 3 per 4 096 bits of text) are queued and validated TOGETHER, one per lane
// (validate_starts): the code lengths decode without error, the lit/len code is complete, the distance
// code complete or a lone code, end-of-block coded.  Until the third session of round 2 every survivor was
// validated on its own, its ~300 code lengths decoded by one lane: 45 % of k_split_find's instructions at
// one active thread (profiles/r2_split_ncu.txt).  Returns the bit position or 0xffffffff.  A false positive is
// caught later: the previous chunk's decoder must land exactly on this bit.
// ------------------------------------------------------------------------------------------------
constexpr uint32_t FIND_KRAFT_BYTES = 1024, FIND_LUT_BYTES = 128 * 32, FIND_QUEUE = 64;
static_assert(sizeof(WSmem) >= FIND_KRAFT_BYTES + FIND_LUT_BYTES + 2 * FIND_QUEUE * 4, "the search's tables live in the decoder's shared memory");

// queue[0, n), n <= 32, ascending: the first one that is a valid dynamic-block header, or 0xffffffff.
// lutb: [128][32] bytes, entry e of lane l at lutb[32 e + l]: symbol | code length << 5 of the code-length code.
__device__ inline uint32_t validate_starts(const In &in, const uint32_t *queue, uint32_t n, uint8_t *lutb, int lane) {
  const bool act = (uint32_t)lane < n;
  const uint32_t q = act ? queue[lane] : 0u;
  uint32_t total = 0, hlit = 0, pp = 0;
  unsigned long long clp = 0;                   // the code-length code's lengths, 3 bits per symbol
  unsigned long long cntp = 0;                  // symbols per length, 8 bits per length
  if (act) {
    const uint32_t v = peek32(in, q + 3);
    hlit = (v & 31u) + 257u;
    const uint32_t hdist = ((v >> 5) & 31u) + 1u, ncl = ((v >> 10) & 15u) + 4u;
    total = hlit + hdist;
    const unsigned long long bits = (unsigned long long)peek32(in, q + 17) | ((unsigned long long)peek32(in, q + 49) << 32);
#pragma unroll
    for (int i = 0; i < 19; i++) {
      const uint32_t l = (uint32_t)i < ncl ? (uint32_t)(bits >> (3 * i)) & 7u : 0u;
      clp |= (unsigned long long)l << (3 * c_clen_order[i]);
      cntp += 1ull << (8 * l);
    }
    pp = q + 17u + 3u * ncl;
  }
  unsigned long long nxp = 0;                   // next canonical code per length, 8 bits per length
  {
    uint32_t code = 0;
#pragma unroll
    for (int L = 1; L <= 7; L++) {
      code = (code + (L > 1 ? (uint32_t)(cntp >> (8 * (L - 1))) & 0xffu : 0u)) << 1;
      nxp |= (unsigned long long)(code & 0xffu) << (8 * L);
    }
  }
  // ---- every lane's own 7-bit table (the code is complete: the filter checked its Kraft sum)
#pragma unroll 1
  for (int sym = 0; sym < 19; sym++) {
    const uint32_t l = act ? (uint32_t)(clp >> (3 * sym)) & 7u : 0u;
    const uint32_t c = (uint32_t)(nxp >> (8 * l)) & 0xffu;
    if (l) nxp += 1ull << (8 * l);
    uint32_t e = l ? __brev(c) >> (32u - l) : 128u;
    const uint32_t fills = __reduce_max_sync(TBZ_FULL, l ? 128u >> l : 0u);
    for (uint32_t it = 0; it < fills; it++) {
      if (e < 128u) { lutb[32u * e + lane] = (uint8_t)(sym | (l << 5)); e += 1u << l; }
    }
  }
  __syncwarp();
  // ---- the code lengths of the two codes, with their Kraft sums (units of 2^-15) as they arrive
  uint32_t idx = 0, last = 0xffu, kr_ll = 0, kr_d = 0, ndist = 0, dlen1 = 0, eob = 0, nlen = 0;
  bool run = act, err = false;
  while (__any_sync(TBZ_FULL, run)) {
    if (run) {
      const uint32_t w = peek32(in, pp);
      const uint32_t ent = lutb[32u * (w & 127u) + lane];
      const uint32_t L = ent >> 5, sym = ent & 31u;
      const uint32_t xb = sym < 16u ? 0u : sym == 16u ? 2u : sym == 17u ? 3u : 7u;
      const uint32_t extra = (w >> L) & ((1u << xb) - 1u);
      uint32_t rep, val;
      if (sym < 16u) { rep = 1u; val = sym; last = sym; }
      else if (sym == 16u) { if (last >= 16u) err = true; rep = 3u + extra; val = last & 15u; }
      else { rep = (sym == 17u ? 3u : 11u) + extra; val = 0u; last = 0u; }
      if (L == 0u || pp + L + xb > in.end || idx + rep > total) err = true;
      pp += L + xb;
      if (val && !err) {
        const uint32_t in_ll = idx + rep <= hlit ? rep : idx < hlit ? hlit - idx : 0u;   // (a run may cross from one code into the other)
        kr_ll += in_ll * (32768u >> val);
        kr_d += (rep - in_ll) * (32768u >> val);
        if (rep > in_ll) { ndist += rep - in_ll; dlen1 = val; }
        if (idx + in_ll > 257u) nlen += idx + in_ll - max(idx, 257u);       // length symbols (257 ...) that have a code
        if (idx <= 256u && 256u < idx + rep) eob = val;
        if (kr_ll > 32768u || kr_d > 32768u) err = true;           // over-subscribed: cannot become complete any more
      }
      idx += rep;
      if (err || idx >= total) run = false;
    }
  }
  // A block must be able to end; lit/len complete; distances complete — or, what an encoder may write for a block with
  // one distance or none: a lone 1-bit code, or no code at all in a block whose lit/len code has no length symbols.
  // (The distance test is where the false positives of the search get through — five on a 1 GiB member before this
  // rule, HDIST 1 .. 7 each, gpurun_out/r2fp.log; everything the search rejects wrongly only costs parallelism.)
  const bool valid = act && !err && idx == total && eob != 0u && kr_ll == 32768u &&
                     (kr_d == 32768u || (ndist == 1u && dlen1 == 1u) || (ndist == 0u && nlen == 0u));
  const uint32_t m = __ballot_sync(TBZ_FULL, valid);
  __syncwarp();
  return m ? queue[__ffs(m) - 1] : 0xffffffffu;
}

__device__ inline uint32_t find_block_start(const In &in, uint32_t from, uint32_t to, WSmem &sm, int lane) {
  uint8_t *const raw = reinterpret_cast<uint8_t *>(&sm);
  // Kraft sums of three 3-bit code lengths at a time (the decoder's tables are free during a search)
  uint16_t *const kraft = reinterpret_cast<uint16_t *>(raw);
  uint8_t *const lutb = raw + FIND_KRAFT_BYTES;
  uint32_t *const queue = reinterpret_cast<uint32_t *>(raw + FIND_KRAFT_BYTES + FIND_LUT_BYTES);   // survivors of both tests
  uint32_t *const q1 = queue + FIND_QUEUE;                                                         // survivors of the first
  __syncwarp();
  for (uint32_t x = lane; x < 512u; x += 32u) {
    uint32_t sum3 = 0;
#pragma unroll
    for (int k = 0; k < 3; k++) { const uint32_t l = (x >> (3 * k)) & 7u; if (l) sum3 += 128u >> l; }
    kraft[x] = (uint16_t)sum3;
  }
  __syncwarp();
  uint32_t qn = 0, q1n = 0;                            // queued candidates (uniform)
  for (uint32_t base = from; base < to || q1n; base += 32) {
    // ---- first test, every lane its own bit position: block type, HLIT and HDIST in range (22 % pass)
    if (base < to) {
      const uint32_t p = base + lane;
      bool c1 = p < to && p + 17 + 12 <= in.end;
      if (c1) {
        const uint32_t h = peek32(in, p);
        // (not the final block: there is one per stream, at its end, and half of the false positives claim to be it)
        c1 = (h & 7u) == 4u && ((h >> 3) & 31) <= 29 && ((h >> 8) & 31) <= 29;
      }
      const uint32_t m1 = __ballot_sync(TBZ_FULL, c1);
      if (c1) q1[q1n + __popc(m1 & ((1u << lane) - 1u))] = p;
      q1n += __popc(m1);
      __syncwarp();
      if (q1n < 32u && base + 32u < to) continue;      // (collect a warp's worth of them for the second test)
    }
    // ---- second test, one queued position per lane: the Kraft sum of the code-length code, 3 ncl <= 57 bits
    const uint32_t take = min(32u, q1n);
    bool cand = (uint32_t)lane < take;
    const uint32_t p = cand ? q1[lane] : 0u;
    const uint32_t moved = (uint32_t)lane + 32u < q1n ? q1[lane + 32u] : 0u;
    if (cand) {
      const uint32_t h = peek32(in, p);
      const uint32_t ncl = ((h >> 13) & 15) + 4, nb = 3u * ncl, q = p + 17;
      uint32_t lo = peek32(in, q), hi = peek32(in, q + 32);
      if (nb < 32u) { lo &= (1u << nb) - 1u; hi = 0u; } else hi &= (1u << (nb - 32u)) - 1u;
      const uint32_t sum = (uint32_t)kraft[lo & 511u] + kraft[(lo >> 9) & 511u] + kraft[(lo >> 18) & 511u] +
                           kraft[((lo >> 27) | (hi << 5)) & 511u] + kraft[(hi >> 4) & 511u] + kraft[(hi >> 13) & 511u] +
                           kraft[(hi >> 22) & 511u];
      cand = sum == 128u && p + 17 + nb <= in.end;
    }
    __syncwarp();
    if ((uint32_t)lane + 32u < q1n) q1[lane] = moved;  // what the queue holds beyond this batch moves to its front
    q1n -= take;
    const uint32_t m = __ballot_sync(TBZ_FULL, cand);
    if (m) {
      if (cand) queue[qn + __popc(m & ((1u << lane) - 1u))] = p;
      qn += __popc(m);
      __syncwarp();
      if (qn > FIND_QUEUE - 32u) {                     // (rare: the queue could not take another batch's survivors)
        for (uint32_t k = 0; k < qn; k += 32u) {
          const uint32_t r = validate_starts(in, queue + k, min(32u, qn - k), lutb, lane);
          if (r != 0xffffffffu) return r;
        }
        qn = 0;
      }
    }
    __syncwarp();
  }
  for (uint32_t k = 0; k < qn; k += 32u) {
    const uint32_t r = validate_starts(in, queue + k, min(32u, qn - k), lutb, lane);
    if (r != 0xffffffffu) return r;
  }
  return 0xffffffffu;
}

}  // namespace tbzfast
