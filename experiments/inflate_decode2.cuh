// inflate_decode2.cuh — phase one of the batched fast path, round-2 decoder: Huffman decode into a stream of
// 64-bit tokens.  Same algorithm as inflate_decode.cuh (one warp per member, 32 speculative sub-chunks per round,
// self-synchronisation proved against checkpoints; that file documents it and still serves the split decode of one
// large member) — rebuilt around the instruction count, which is what bounds this kernel (ALU pipe at 63 %,
// round 1: ~140 thread-instructions per symbol pair):
//   * table entries are 32 bits and carry everything a symbol needs — code length in the low five bits (so the
//     entry itself is the shift amount), extra-bit count, the BASE VALUE (literal byte, length - 3, distance - 1)
//     and the total number of bits to drop: no base-value tables, no selects on the symbol kind
//     (replaces the 16-bit nodes + constants.lisp:36-61 lookups of the reference, huffman-tree.lisp:15-76)
//   * the bit reader keeps two stream words and a bit offset; a peek is one funnel shift, dropping bits is one add,
//     a refill is three moves and a load (deflate.lisp:142-231 keeps a shifted 64-bit accumulator instead)
//   * a token is `up to four literals + one match` (8 bytes): phase two then has exactly one back-reference per
//     lane and step; literals are gathered in a register until a match closes the token
//   * no output offsets are tracked here (phase two scans the lengths anyway; it also owns the overflow verdict)
// Codes longer than the root tables (10 bits lit/len, 9 bits distance: ~1 % of the symbols on text) take a canonical
// search.  Shared memory is what limits the number of members in flight: 7.8 KB per warp, 28 warps per SM.  (A first
// version kept the sorted symbol lists of that search in global memory: with all of shared memory taken L1 is 4 KB,
// every search was an L2 round trip, and half of all loop iterations have a lane that searches: 0.79 ms against 0.62.)
#pragma once
#include "tbz_device.cuh"
#include "inflate_decode.cuh"

namespace tbzd2 {

#if defined(TBZ_EMU) && defined(TBZ_EMU_TRACE)
#define TBZ_D2_WHY(what) do { if (lane == 0) fprintf(stderr, "[d2] %s (line %d)\n", what, __LINE__); } while (0)
#else
#define TBZ_D2_WHY(what) do { } while (0)
#endif

using tbzfast::byte_at;
using tbzfast::Canon16;
using tbzfast::canon_lookup;
using tbzfast::In;
using tbzfast::ldw;
using tbzfast::member_start;
using tbzfast::NL;
using tbzfast::NO_SLAB;
using tbzfast::P1Rec;
using tbzfast::peek32;
using tbzfast::SLAB_HDR_WORDS;
using tbzfast::SLAB_WORDS;
using tbzfast::SlabHdr;
using tbzfast::warp_canon;

constexpr int WPC = 4;                              // warps (members in flight) per CTA
constexpr int NT = WPC * 32;
constexpr int KLL = 10, KD = 9;                     // root table bits
constexpr uint32_t TOKCAP2 = tbzfast::TOKCAP / 2;   // 64-bit tokens a list has room for (same slab geometry as the 32-bit lists)
constexpr uint32_t LANECAP = 256;                   // tokens a lane emits per round before it ends the round early
constexpr uint32_t CKSTEP2 = 16;                    // a checkpoint about every 16 tokens
constexpr uint32_t NCK2 = LANECAP / CKSTEP2;
constexpr uint32_t S_MAX = 4000, S_MIN = 64;        // sub-chunk size in bits (12-bit field in a checkpoint: < 4095)
constexpr uint16_t CK_NONE = 0xffffu;
static_assert(tbzfast::TOKCAP % 2 == 0 && LANECAP % CKSTEP2 == 0 && LANECAP <= TOKCAP2 && S_MAX < 4095, "token list geometry");

// ---- 64-bit token: lo = up to four literal bytes (first byte lowest); hi: [7:0] match length - 3, [22:8] distance - 1,
//      [25:23] number of literals, bit 31 = a match follows the literals
constexpr uint32_t T2_MATCH = 0x80000000u;
__device__ __forceinline__ uint32_t t2_nlit(uint32_t hi) { return (hi >> 23) & 7u; }
__device__ __forceinline__ uint32_t t2_outlen(uint32_t hi) { return t2_nlit(hi) + ((hi & T2_MATCH) ? (hi & 255u) + 3u : 0u); }

// ---- table entry: [4:0] code length, [7:5] kind, [22:8] base value, [26:23] extra bits, [31:27] bits to drop (length + extra)
constexpr uint32_t K_LIT = 0u << 5, K_LEN = 1u << 5, K_EOB = 2u << 5, K_LONG = 3u << 5, K_INVALID = 4u << 5, K_MASK = 7u << 5;
constexpr uint32_t K_SPECIAL = 6u << 5;             // any of these bits: end of block, long code, no code
__device__ __forceinline__ uint32_t mk_entry(uint32_t kind, uint32_t L, uint32_t base, uint32_t xb) {
  return L | kind | (base << 8) | (xb << 23) | ((L + xb) << 27);
}
__device__ __forceinline__ uint32_t ll_entry(uint32_t sym, uint32_t L) {
  if (sym < 256) return mk_entry(K_LIT, L, sym, 0);
  if (sym == 256) return mk_entry(K_EOB, L, 0, 0);
  if (sym > 285) return K_INVALID;                                    // huffman-tree.lisp:176-177
  return mk_entry(K_LEN, L, (uint32_t)c_len_base[sym - 257] - 3u, c_len_extra[sym - 257]);
}
__device__ __forceinline__ uint32_t d_entry(uint32_t sym, uint32_t L) {
  if (sym > 29) return K_INVALID;                                     // huffman-tree.lisp:172-175
  return mk_entry(K_LIT, L, (uint32_t)c_dist_base[sym] - 1u, c_dist_extra[sym]);
}

struct HdrScratch {                      // only alive while a block header is parsed and the tables are built
  uint16_t lut_cl[128];
  Canon16 c_cl;
  uint16_t sorted_cl[32];
  uint8_t lens[352];                     // [0,19) code-length code, [32,352) lit/len + distance
  uint16_t run[16];
};
struct WSmem {                           // one per warp
  uint32_t lut[(1 << KLL) + (1 << KD)];  // lit/len root table, the distance root table directly behind it
  union {
    uint16_t ckpt[NCK2][NL];             // [checkpoint][lane]: bit offset in the sub-chunk | (token index - 16 c) << 12
    HdrScratch h;
  };
  Canon16 c_ll, c_d;                     // first code / count / base per length, and the symbols sorted by (length, symbol):
  uint16_t sorted_ll[288], sorted_d[32]; // the search for codes beyond the root tables (~1 % of the symbols on text)
};
static_assert(sizeof(HdrScratch) <= sizeof(uint16_t) * NCK2 * NL, "header scratch must fit under the checkpoints");
static_assert((sizeof(WSmem) * WPC + 1024) * 7 <= 232448, "seven CTAs = 28 warps per SM");

// ---- bit reader: two stream words + the next one prefetched; bo < 32 between calls
struct BR { uint32_t w0, w1, nw, wi, bo; };
__device__ __forceinline__ void br_init(BR &b, const In &in, uint32_t pos) {
  const uint32_t q = pos >> 5;
  b.w0 = ldw(in, q); b.w1 = ldw(in, q + 1); b.nw = ldw(in, q + 2); b.wi = q + 3; b.bo = pos & 31u;
}
__device__ __forceinline__ uint32_t br_peek(const BR &b) { return __funnelshift_r(b.w0, b.w1, b.bo); }   // 32 valid bits
__device__ __forceinline__ void br_skip(BR &b, const In &in, uint32_t n) {                               // n <= 31
  b.bo += n;
  if (b.bo >= 32u) { b.w0 = b.w1; b.w1 = b.nw; b.nw = ldw(in, b.wi); b.wi++; b.bo -= 32u; }
}

// ---- token list writer: one 8-byte store per token.  (Pairing two tokens into a 16-byte store cost ~20 register
// moves per token in this loop; global stores are ~4 % of the LSU wavefronts here, so the wider store buys nothing.)
struct TokW {
  uint2 *wp;
  __device__ __forceinline__ void open(uint2 *list) { wp = list; }
  __device__ __forceinline__ void emit(uint32_t lo, uint32_t hi) { *wp = make_uint2(lo, hi); wp++; }
};

struct Lane {                            // decode state of one lane within a round
  BR b;
  uint32_t p;                            // bit position of the next symbol
  uint32_t k;                            // tokens emitted
  uint32_t lb, nl;                       // literals waiting for their match
  TokW tw;
};

__device__ __forceinline__ void flush_literals(Lane &s) {
  if (s.nl) { s.tw.emit(s.lb, s.nl << 23); s.k++; s.lb = 0; s.nl = 0; }
}

// One item: a literal (two, if the next symbol is a literal as well) or a length + distance pair.  Every lane runs the
// same instructions for either.  Returns 0 = go on, 2 = end of block (its bits dropped), 3 = no such code.
// stop_at: a bit position a second literal must not start at... it ends the item instead (phase 1b: the place where
// this lane may synchronise with the next one has to be the start of an item, whatever the pairing was so far).
__device__ __forceinline__ int item(Lane &s, const In &in, const WSmem &sm, uint32_t stop_at) {
  const uint32_t w = br_peek(s.b);
  uint32_t e = sm.lut[w & ((1u << KLL) - 1u)];
  if (__builtin_expect((e & K_SPECIAL) != 0, 0)) {
    if ((e & K_MASK) == K_LONG) {
      const uint32_t r = canon_lookup(sm.c_ll, sm.sorted_ll, w, KLL + 1, 15);
      e = r ? ll_entry(r >> 4, r & 15u) : K_INVALID;
    }
    if ((e & K_MASK) == K_EOB) { s.p += e & 31u; return 2; }
    if (e & K_SPECIAL) return 3;
  }
  const uint32_t v1 = ((e >> 8) & 0x7fffu) + (__funnelshift_r(w, 0u, e) & ~(0xffffffffu << ((e >> 23) & 15u)));
  const uint32_t n1 = e >> 27;
  br_skip(s.b, in, n1);
  const bool ism = (e & K_LEN) != 0;
  const uint32_t w2 = br_peek(s.b);
  uint32_t e2 = sm.lut[ism ? (1u << KLL) + (w2 & ((1u << KD) - 1u)) : (w2 & ((1u << KLL) - 1u))];
  if (__builtin_expect(ism && (e2 & K_SPECIAL) != 0, 0)) {
    if ((e2 & K_MASK) == K_LONG) {
      const uint32_t r = canon_lookup(sm.c_d, sm.sorted_d, w2, KD + 1, 15);
      e2 = r ? d_entry(r >> 4, r & 15u) : K_INVALID;
    }
    if (e2 & K_SPECIAL) return 3;
  }
  const bool two = !ism && (e2 & K_MASK) == K_LIT && s.p + n1 != stop_at;   // a second literal travels with the first
  const uint32_t v2 = ((e2 >> 8) & 0x7fffu) + (__funnelshift_r(w2, 0u, e2) & ~(0xffffffffu << ((e2 >> 23) & 15u)));
  const uint32_t n2 = (ism || two) ? e2 >> 27 : 0u;
  br_skip(s.b, in, n2);
  s.p += n1 + n2;
  // one place where a token leaves: a match closes it; literals close it when it would hold more than four
  const uint32_t cnt = two ? 2u : 1u;
  if (ism || s.nl + cnt > 4u) {
    s.tw.emit(s.lb, (s.nl << 23) | (ism ? T2_MATCH | (v2 << 8) | v1 : 0u));
    s.k++; s.lb = 0; s.nl = 0;
  }
  if (!ism) {
    s.lb |= (two ? v1 | (v2 << 8) : v1) << (8u * s.nl);
    s.nl += cnt;
  }
  return 0;
}

enum { ST_RUN = 0, ST_OVER, ST_END, ST_SYNC, ST_EOB, ST_CAP, ST_BAD, ST_IDLE };   // ST_IDLE: the lane had nothing to decode

// ------------------------------------------------------------------------------------------------
// Every block of a member from bit `pos` on, one warp.  Returns true when the token stream is complete (rec filled in),
// false when the member goes to the sequential kernel.  Every return value is warp-uniform.
// ------------------------------------------------------------------------------------------------
__device__ inline bool decode_blocks(const In &in, uint32_t pos, P1Rec &rec, WSmem &sm,
                                     uint32_t *__restrict__ slabs, uint32_t nslabs, uint32_t *slab_counter, int lane) {
  uint32_t first_slab = NO_SLAB, prev_slab = NO_SLAB;
  uint32_t prev_block_bits = 0;   // size of the previous block of this member: predicts this one
  bool last = false;
  uint16_t *const sorted_ll = sm.sorted_ll, *const sorted_d = sm.sorted_d;

  while (!last) {
    // ================= block header (deflate.lisp:518-528, :577-669) =================
    if (in.end - pos < 3) { TBZ_D2_WHY("give up"); return false; }
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
      if (in.end - pos < 14) { TBZ_D2_WHY("give up"); return false; }
      const uint32_t v = peek32(in, pos);
      hlit = (v & 31) + 257; hdist = ((v >> 5) & 31) + 1;
      const int ncl = ((v >> 10) & 15) + 4;
      if (in.end - pos < 14u + 3u * ncl) { TBZ_D2_WHY("give up"); return false; }
      if (lane < 19) sm.h.lens[lane] = 0;
      __syncwarp();
      if (lane < ncl) sm.h.lens[c_clen_order[lane]] = peek32(in, pos + 14 + 3 * lane) & 7;
      __syncwarp();
      int err = warp_canon(sm.h.lens, 19, sm.h.c_cl, sm.h.sorted_cl, sm.h.run, lane);
      if (!err && sm.h.c_cl.nsyms == 0) err = TBZ_ERR_INVALID_SYMBOL;
      if (err) { TBZ_D2_WHY("give up"); return false; }
      // entry: [3:0] code length, [7:4] extra bits, [12:8] symbol; 0 = no code
      for (int e = lane; e < 128; e += 32) {
        const uint32_t r = canon_lookup(sm.h.c_cl, sm.h.sorted_cl, (uint32_t)e, 1, 7);
        const uint32_t sym = r >> 4;
        const uint32_t xb = sym < 16 ? 0 : sym == 16 ? 2 : sym == 17 ? 3 : 7;
        sm.h.lut_cl[e] = (uint16_t)(r ? ((r & 15) | (xb << 4) | (sym << 8)) : 0);
      }
      __syncwarp();
      uint32_t p = pos + 14 + 3 * ncl;
      if (lane == 0) {
        // the code lengths themselves: one lane, table driven (deflate.lisp:626-669)
        int idx = 0, lastlen = 0xff;
        const int total = hlit + hdist;
        BR hb;
        br_init(hb, in, p);
        while (idx < total) {
          const uint32_t w = br_peek(hb);
          const uint32_t r = sm.h.lut_cl[w & 127];
          if (!r) { err = 1; break; }
          const uint32_t L = r & 15, xb = (r >> 4) & 15, sym = r >> 8;
          if (p + L + xb > in.end) { err = 1; break; }
          p += L + xb;
          br_skip(hb, in, L + xb);
          if (sym < 16) { sm.h.lens[32 + idx] = (uint8_t)sym; idx++; lastlen = (int)sym; continue; }
          const uint32_t extra = (w >> L) & ((1u << xb) - 1);
          int rep, val;
          if (sym == 16) { if (lastlen >= 16) { err = 1; break; } rep = 3 + extra; val = lastlen; }
          else { rep = (sym == 17 ? 3 : 11) + extra; val = 0; lastlen = 0; }
          if (idx + rep > total) { err = 1; break; }
          for (int q = 0; q < rep; q++) sm.h.lens[32 + idx + q] = (uint8_t)val;
          idx += rep;
        }
      }
      err = __shfl_sync(TBZ_FULL, err, 0);
      if (err) { TBZ_D2_WHY("give up"); return false; }
      pos = __shfl_sync(TBZ_FULL, p, 0);
    } else if (btype == 0) {
      // ================= stored block (deflate.lisp:532-573): LEN, NLEN, then LEN bytes as they are =================
      // They travel as literal tokens, four bytes each, the lanes taking consecutive slices.
      pos = (pos + 7u) & ~7u;
      if (in.end < pos || in.end - pos < 32u) { TBZ_D2_WHY("give up"); return false; }
      const uint32_t v = peek32(in, pos);
      uint32_t slen = v & 0xffffu;
      if ((slen ^ 0xffffu) != (v >> 16)) { TBZ_D2_WHY("give up"); return false; }                 // deflate.lisp:535
      pos += 32u;
      if (in.end - pos < 8u * slen) { TBZ_D2_WHY("give up"); return false; }                      // the input ends inside the block
      uint32_t bp = pos >> 3;                                          // byte offset of the payload from in.w
      pos += 8u * slen;
      while (slen) {
        uint32_t slab_id = 0;
        if (lane == 0) slab_id = atomicAdd(slab_counter, 1u);
        slab_id = __shfl_sync(TBZ_FULL, slab_id, 0);
        if (slab_id >= nslabs) { TBZ_D2_WHY("give up"); return false; }
        uint32_t *slab = slabs + (size_t)slab_id * SLAB_WORDS;
        SlabHdr *sh = reinterpret_cast<SlabHdr *>(slab);
        const uint32_t nb = slen < (uint32_t)(NL * TOKCAP2 * 4) ? slen : (uint32_t)(NL * TOKCAP2 * 4);
        const uint32_t per = ((nb + NL - 1) / NL + 3u) & ~3u;          // bytes per lane, a multiple of four
        const uint32_t lo = per * lane < nb ? per * lane : nb, hi = lo + per < nb ? lo + per : nb;
        const uint32_t cnt = (hi - lo + 3u) / 4u;
        uint2 *list = reinterpret_cast<uint2 *>(slab + SLAB_HDR_WORDS) + lane * TOKCAP2;
        for (uint32_t t = 0; t < cnt; t++) {
          const uint32_t a = lo + 4u * t, m = hi - a < 4u ? hi - a : 4u;
          uint32_t wv = peek32(in, (bp + a) * 8u);
          if (m < 4u) wv &= (1u << (8u * m)) - 1u;
          list[t] = make_uint2(wv, m << 23);
        }
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
      continue;
    } else {
      { TBZ_D2_WHY("give up"); return false; }                        // reserved block type: sequential kernel
    }
    __syncwarp();
    // ================= tables (huffman-tree.lisp:99-218) =================
    if (warp_canon(sm.h.lens + 32, hlit, sm.c_ll, sorted_ll, sm.h.run, lane)) { TBZ_D2_WHY("give up"); return false; }
    if (warp_canon(sm.h.lens + 32 + hlit, hdist, sm.c_d, sorted_d, sm.h.run, lane)) { TBZ_D2_WHY("give up"); return false; }
    if (sm.c_ll.nsyms == 0) { TBZ_D2_WHY("give up"); return false; }
    for (int e = lane; e < (1 << KLL); e += 32) {
      const uint32_t r = canon_lookup(sm.c_ll, sorted_ll, (uint32_t)e, 1, KLL);
      sm.lut[e] = r ? ll_entry(r >> 4, r & 15) : (sm.c_ll.maxlen > KLL ? K_LONG : K_INVALID);
    }
    for (int e = lane; e < (1 << KD); e += 32) {
      const uint32_t r = canon_lookup(sm.c_d, sorted_d, (uint32_t)e, 1, KD);
      sm.lut[(1 << KLL) + e] = r ? d_entry(r >> 4, r & 15) : (sm.c_d.maxlen > KD ? K_LONG : K_INVALID);
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
      if (slab_id >= nslabs) { TBZ_D2_WHY("give up"); return false; }
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
      for (uint32_t c = 0; c < NCK2; c++) sm.ckpt[c][lane] = CK_NONE;
      uint32_t *slab = slabs + (size_t)slab_id * SLAB_WORDS;
      SlabHdr *sh = reinterpret_cast<SlabHdr *>(slab);
      Lane s;
      s.tw.open(reinterpret_cast<uint2 *>(slab + SLAB_HDR_WORDS) + lane * TOKCAP2);
      s.k = 0; s.lb = 0; s.nl = 0;
      __syncwarp();

      // ---- 1a: speculative decode of the lane's sub-chunk.  The loop is kept warp-converged (one vote per
      // iteration, the body under `act`).  A checkpoint is taken about every CKSTEP2 tokens: the token list is
      // complete up to this bit (literals that wait for a match close their token first).
      const uint32_t cstart = P0 + S * lane, cend = cstart + S;
      s.p = cstart;
      int st = s.p < in.end ? ST_RUN : ST_IDLE;
      if (st == ST_RUN) br_init(s.b, in, s.p);
      uint32_t nextck = 0;
      while (__any_sync(TBZ_FULL, st == ST_RUN)) {
        if (st == ST_RUN) {
          if (s.p >= cend) st = ST_OVER;
          else if (s.k + 2u >= LANECAP) st = ST_CAP;
          else {
            if (s.k >= nextck) {                                   // a checkpoint: literals still waiting close their token here
              flush_literals(s);
              if (s.k - nextck < 16u) sm.ckpt[nextck / CKSTEP2][lane] = (uint16_t)((s.p - cstart) | ((s.k - nextck) << 12));
              nextck = (s.k & ~(CKSTEP2 - 1u)) + CKSTEP2;
            }
            const int r = item(s, in, sm, 0xffffffffu);
            if (r) st = r == 2 ? ST_EOB : ST_BAD;
          }
        }
      }
      __syncwarp();
      // ---- 1b: past the own sub-chunk: decode on until the start of an item coincides with a checkpoint
      // of the lane whose sub-chunk the position lies in.  Same converged loop; an iteration either
      // decodes one item or looks up the next place a synchronisation can happen (`tgt`).
      uint32_t nx = 0, g_sync = 0;
      {
        uint32_t j = lane, jend = cend, c = 0, tgt = 0;
        while (__any_sync(TBZ_FULL, st == ST_OVER)) {
          if (st == ST_OVER) {
            if (s.p >= tgt) {
              if (s.p >= winend) st = ST_END;
              else {
                while (s.p >= jend) { j++; jend += S; c = 0; }
                const uint32_t rel = s.p - (jend - S);
                uint32_t ck = CK_NONE;
                while (c < NCK2 && ((ck = sm.ckpt[c][j]) == CK_NONE || (ck & 0xfffu) < rel)) c++;
                if (c < NCK2 && (ck & 0xfffu) == rel) {
                  st = ST_SYNC; nx = j; g_sync = c * CKSTEP2 + (ck >> 12);
                } else {
                  tgt = c < NCK2 ? (jend - S) + (ck & 0xfffu) : jend;
                }
              }
            } else if (s.k + 2u >= LANECAP) st = ST_CAP;
            else {
              const int r = item(s, in, sm, tgt);
              if (r) st = r == 2 ? ST_EOB : ST_BAD;
            }
          }
        }
      }
      // a lane that decoded past the end of the input has nothing proven to offer (zeros are read there)
      if (st != ST_IDLE && s.p > in.end) st = ST_BAD;
      if (st != ST_BAD) flush_literals(s);            // the list ends exactly at bit s.p
      __syncwarp();
      // ---- 1c: lanes reachable from lane 0 through "synchronised into" edges are proven
      uint32_t my_g = 0;
      bool proven = false;
      int term_st;
      uint32_t term_pos;
      {
        int cur = 0;
        for (;;) {
          if (lane == cur) proven = true;
          const int st_c = __shfl_sync(TBZ_FULL, st, cur);
          if (st_c != ST_SYNC) { term_st = st_c; term_pos = __shfl_sync(TBZ_FULL, s.p, cur); break; }
          const uint32_t nx_c = __shfl_sync(TBZ_FULL, nx, cur);
          const uint32_t g_c = __shfl_sync(TBZ_FULL, g_sync, cur);
          if ((uint32_t)lane == nx_c) my_g = g_c;
          cur = (int)nx_c;
        }
      }
      if (term_st == ST_BAD || term_st == ST_IDLE || term_st == ST_OVER) { TBZ_D2_WHY("give up"); return false; }
      // a lane that ran into its token cap ends the round early: use shorter sub-chunks from here on
      if (term_st == ST_CAP && shrink < 6) shrink++;
      const uint32_t cnt = proven ? s.k - my_g : 0u;
      sh->fc[lane] = cnt ? (my_g | (cnt << 16)) : 0u;
      if (lane == 0) {
        sh->next = NO_SLAB; sh->out_bytes = 0;
        if (prev_slab != NO_SLAB) reinterpret_cast<SlabHdr *>(slabs + (size_t)prev_slab * SLAB_WORDS)->next = slab_id;
      }
      if (first_slab == NO_SLAB) first_slab = slab_id;
      prev_slab = slab_id;
#if defined(TBZ_EMU) && defined(TBZ_EMU_TRACE)
      if (lane == 0) fprintf(stderr, "[d2] round P0 %u S %u -> term_st %d pos %u (slab %u)\n", P0, S, term_st, term_pos, slab_id);
#endif
      // ---- how did the round end?
      if (term_pos <= pos && term_st != ST_EOB) { TBZ_D2_WHY("give up"); return false; }     // no progress (cannot happen; guards the loop)
      pos = term_pos;
      if (term_st == ST_EOB) block_done = true;
      __syncwarp();
    }
    prev_block_bits = pos - data_start;
  }
  if (lane == 0) {
    rec.first_slab = first_slab;
    rec.out_len = 0xffffffffu;             // not tracked here: phase two counts (and owns the overflow verdict)
    rec.end_pos = pos;
    rec.status = 1u;
  }
  return true;
}

// One member, one warp: wrapper header, then every block.
__device__ inline bool decode_member(const DMember &mem, int fmt, P1Rec &rec, WSmem &sm,
                                     uint32_t *__restrict__ slabs, uint32_t nslabs, uint32_t *slab_counter, int lane) {
  In in;
  uint32_t pos;
  if (!member_start(mem, fmt, in, pos)) { TBZ_D2_WHY("give up"); return false; }
  return decode_blocks(in, pos, rec, sm, slabs, nslabs, slab_counter, lane);
}

}  // namespace tbzd2
