// huff_decode.cuh — phase one of the batched fast path (round 2): Huffman decode of a member into token lists,
// ONE WARP per member, every lane a decode lane.  It replaces the decode loop of round 1 (inflate_decode.cuh, which
// still serves the split decode of one large member and documents the scheme) and writes the same slabs of 32-bit
// tokens, so phase two (inflate_copy.cuh) is unchanged.
//
// The compressed bits of a block are cut into 32 sub-chunks of S bits.  Lane i starts at the first bit of
// sub-chunk i without knowing whether a symbol starts there (only lane 0 does) and relies on the
// self-synchronisation of Huffman streams (117 bits on average on the text of BASELINE config 2, p99 595):
//   * every lane decodes into its own token list (global memory) and records where its tokens start — every fourth
//     one at first (shared memory), then every sixteenth (global scratch, rarely looked at); all lanes emit one
//     token per iteration, so the iteration count is the token index and the records are taken by the whole warp
//     at the same iterations
//   * a lane does not stop at the end of its sub-chunk: it decodes on until one of its tokens starts where the lane
//     that owns those bits recorded a token start — from there on both decodes are identical (same tables, same
//     bit), so the owner's list is proven from that token on and this lane is done.  One loop does both
//   * a walk over "who synchronised into whom" from lane 0 gives the proven lanes and their first proven token
// What this kernel is bound by is its instruction count (integer work: the ALU pipe issues a warp instruction every
// other cycle), so the loop body is short and branch-free for the common symbols:
//   * 32-bit table entries carry everything a symbol needs: byte 0 = code length | kind flags, byte 1 = bits to
//     drop (length + extra bits), bits 31..17 = the BASE VALUE (literal, length - 3, distance - 1); the extra bits
//     are cut out of the peeked word with the two counts (replaces the 16-bit nodes and the constants.lisp:36-61
//     lookups of the reference, huffman-tree.lisp:15-76; round 1 had 16-bit entries + two base-value tables)
//   * root tables of 10 (lit/len) and 8 (distance) bits; longer codes go through second-level tables (one more
//     load, no search): the entry of their prefix links to a sub-table indexed by the following bits
//   * an iteration decodes one token with exactly two lookups, the same instructions for every lane: a length and
//     its distance, or a literal and — if the next symbol is a literal too — that one as well
//   * the bit reader keeps three stream words and a bit offset: a peek is a funnel shift, dropping bits an add,
//     one refill point per iteration (deflate.lisp:142-231 keeps a shifted 64-bit accumulator instead).  The stream
//     reaches every lane in 16-byte chunks, two of them waiting or on their way in shared-memory slots with cp.async
//     (LDGSTS): a load into a register that the loop carries forward costs a move that waits for it whatever the
//     distance to its first use, and one 16-byte request per four words is a quarter of the memory pipe's work
//   * four tokens leave with one 16-byte store, at iterations that are the same for the whole warp
// Stored blocks travel as literal tokens.  Output bytes are not counted here: phase two owns the overflow verdict.
// Anything not provably clean (bad codes, truncation, a header this kernel does not take) sends the member to the
// sequential kernel (inflate_seq.cuh), which reproduces the reference's verdict.
// Replaces deflate.lisp:465-509,577-702 (decode) and huffman-tree.lisp:99-218 (tables).
#pragma once
#include "tbz_device.cuh"
#include "inflate_decode.cuh"

namespace tbzhd {

#if defined(TBZ_EMU) && defined(TBZ_EMU_TRACE)
#define TBZ_HD_WHY(what) do { if (lane == 0) fprintf(stderr, "[hd] %s (line %d)\n", what, __LINE__); } while (0)
#else
#define TBZ_HD_WHY(what) do { } while (0)
#endif

using tbzfast::byte_at;
using tbzfast::Canon16;
using tbzfast::canon_lookup;
using tbzfast::In;
using tbzfast::member_start;
using tbzfast::P1Rec;
using tbzfast::peek32;
using tbzfast::warp_canon;

#ifndef TBZ_HD_SMAX
#define TBZ_HD_SMAX 4000
#endif
#ifndef TBZ_HD_KD
#define TBZ_HD_KD 8
#endif
#ifndef TBZ_HD_SUBCAP
#define TBZ_HD_SUBCAP 224
#endif
#ifndef TBZ_HD_MINBLOCKS
#define TBZ_HD_MINBLOCKS 7
#endif
constexpr int WPC = 4;                              // warps (members in flight) per CTA
constexpr int NT = WPC * 32;
constexpr int NL = 32;                              // decode lanes per member
constexpr int KLL = 10, KD = TBZ_HD_KD;             // root table bits
constexpr uint32_t SUBCAP = TBZ_HD_SUBCAP;          // second-level entries a block may need (both codes together)
constexpr uint32_t LUT_D = 1u << KLL, LUT_SUB = LUT_D + (1u << KD), LUT_N = LUT_SUB + SUBCAP;
constexpr uint32_t ITEMCAP = tbzfast::TOKCAP;       // tokens a lane may emit per round (sub-chunk + overrun): the slab geometry of round 1
constexpr uint32_t S_MAX = TBZ_HD_SMAX, S_MIN = 256;  // sub-chunk size in bits
constexpr uint32_t CK_DENSE = 24, CK_SPARSE = 24, NCK = CK_DENSE + CK_SPARSE;   // recorded item starts: items 4 s, s < 24 (shared memory: the bits since the start recorded before), then items 96 + 16 i (global scratch: rarely looked at)
constexpr uint16_t CK_NONE = 0xffffu, CK_END = 0xfffeu;     // not recorded (yet) / the lane records no more
static_assert(S_MAX < 0xfffeu, "a recorded token start is a 16-bit offset into the sub-chunk");

using tbzfast::NO_SLAB;
using tbzfast::SLAB_HDR_WORDS;
using tbzfast::SLAB_WORDS;
using tbzfast::SlabHdr;
constexpr size_t SCRATCH_BYTES = (size_t)NL * (CK_SPARSE + 1) * 4;   // per warp of the grid: the sparse token starts of a round
static_assert(ITEMCAP % 4 == 0 && (SLAB_HDR_WORDS * 4) % 16 == 0, "four tokens per 16-byte store");

// ---- token (what one iteration of the decode loop emits, 32 bits; the format of round 1): a literal = its byte; two
//      literals = I_LIT2 | second << 8 | first; a match = I_MATCH | (distance - 1) << 8 | (length - 3)
constexpr uint32_t I_MATCH = tbzfast::TOK_MATCH, I_LIT2 = tbzfast::TOK_LIT2;

// ---- table entry: [4:0] code length, bit 5 = a length symbol, bit 6 = stop (end of block; with bit 5: no such code),
//      bit 7 = link to a second-level table ([4:0] = its index bits, [31:17] = its offset in the table array);
//      [15:8] bits to drop (code length + extra bits), [31:17] base value
constexpr uint32_t K_LEN = 0x20u, K_STOP = 0x40u, K_SUB = 0x80u, K_SPECIAL = K_STOP | K_SUB;
constexpr uint32_t E_EOB_KIND = K_STOP, E_INVALID = K_STOP | K_LEN;
__device__ __forceinline__ uint32_t mk_entry(uint32_t kind, uint32_t L, uint32_t base, uint32_t xb) {
  return L | kind | ((L + xb) << 8) | (base << 17);
}
__device__ __forceinline__ uint32_t ll_entry(uint32_t sym, uint32_t L) {
  if (sym < 256) return mk_entry(0, L, sym, 0);
  if (sym == 256) return mk_entry(E_EOB_KIND, L, 0, 0);
  if (sym > 285) return E_INVALID;                                    // huffman-tree.lisp:176-177
  return mk_entry(K_LEN, L, (uint32_t)c_len_base[sym - 257] - 3u, c_len_extra[sym - 257]);
}
__device__ __forceinline__ uint32_t d_entry(uint32_t sym, uint32_t L) {
  if (sym > 29) return E_INVALID;                                     // huffman-tree.lisp:172-175
  return mk_entry(0, L, (uint32_t)c_dist_base[sym] - 1u, c_dist_extra[sym]);
}
// value of a decoded symbol: base + the extra bits, which are bits [L, n) of the peeked word
__device__ __forceinline__ uint32_t e_drop(uint32_t e) { return __byte_perm(e, 0u, 0x4441u); }            // byte 1
__device__ __forceinline__ uint32_t e_value(uint32_t e, uint32_t x, uint32_t n) {
  return (e >> 17) + __funnelshift_r(x & ~(0xffffffffu << n), 0u, e);                                     // (shift = e & 31 = L)
}

// ---- shared-space addressing and cp.async for the per-lane input slots
#ifdef TBZ_EMU
__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)((const unsigned char *)p - ::emu::dyn_smem()); }
__device__ __forceinline__ uint32_t lds32(uint32_t a) { return *reinterpret_cast<const uint32_t *>(::emu::dyn_smem() + a); }
__device__ __forceinline__ void cp_async16(uint32_t a, const void *g) { memcpy(::emu::dyn_smem() + a, g, 16); }
__device__ __forceinline__ void cp_async_commit() {}
template <int N> __device__ __forceinline__ void cp_async_wait() {}
#else
__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t lds32(uint32_t a) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ void cp_async16(uint32_t a, const void *g) { asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(a), "l"(g) : "memory"); }
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
#endif

struct HdrScratch {                      // only alive while a block header is parsed and the tables are built
  uint16_t lut_cl[128];
  Canon16 c_cl, c_ll, c_d;
  uint16_t sorted_cl[32], sorted_d[32];
  uint16_t run[16];
};
// The code lengths of a header ([0,19) code-length code, [32,352) lit/len + distance) live where the second-level
// tables will be: they are not needed any more once both codes are sorted.
static_assert(SUBCAP * 4 >= 352, "code lengths under the second-level tables");
struct WSmem {                           // one per warp
  uint32_t lut[LUT_N];                   // lit/len root table, distance root table, second-level tables
  union {
    uint4 inq[2][NL];                    // [slot][lane]: the lane's 16-byte chunk c of the stream waits in slot c % 2 (cp.async landing zone)
    uint16_t sorted_ll[288];             // (while the tables are built: the lit/len symbols sorted by code)
  };
  union {
    uint8_t ckd[CK_DENSE][NL];           // [token][lane]: the bits between the starts of this token and the one before; 0 = not recorded (yet), 1 = the lane records no more
    HdrScratch h;
  };
};
static_assert(sizeof(HdrScratch) <= sizeof(uint8_t) * CK_DENSE * NL, "header scratch must fit under the checkpoints");
static_assert((sizeof(WSmem) * WPC + 1024) * TBZ_HD_MINBLOCKS <= 233472, "CTAs per SM (228 KB, 1 KB of it reserved per CTA)");

// recorded token starts: the dense ones in shared memory (slot = token index; the value is the distance from the start
// recorded before), the sparse ones — offset | token index << 16 — in the warp's global scratch (written by their
// lane, read by another lane of the same warp after a __syncwarp, past L1); the slot behind a sparse one is set to
// "not yet" first
__device__ __forceinline__ void ck_put(WSmem &sm, uint32_t *gck, uint32_t slot, int lane, uint32_t off, uint32_t tok, uint32_t lastoff) {
  if (slot < CK_DENSE) sm.ckd[slot][lane] = (uint8_t)(off - lastoff);
  else {
    uint32_t *g = gck + lane * (CK_SPARSE + 1) + (slot - CK_DENSE);
    g[1] = CK_NONE;
    g[0] = off | (tok << 16);
  }
}
__device__ __forceinline__ void ck_end(WSmem &sm, uint32_t *gck, uint32_t slot, int lane) {
  if (slot < CK_DENSE) sm.ckd[slot][lane] = 1;
  else gck[lane * (CK_SPARSE + 1) + (slot - CK_DENSE)] = CK_END;
}

enum { ST_RUN = 0, ST_END, ST_SYNC, ST_EOB, ST_CAP, ST_BAD, ST_IDLE };   // ST_IDLE: the lane had nothing to decode

// Second-level tables of one code (codes longer than K bits): the symbols sorted by (length, code) — that is the order
// of their left-aligned code values, so the symbols behind one K-bit prefix are neighbours and the last of them is the
// longest.  The last symbol of every prefix sizes that prefix's table and links it into the root table; then every long
// symbol fills its entries.  `sub` (uniform): next free entry of the table array.  false: SUBCAP does not suffice.
template <int K, bool DIST>
__device__ inline bool build_sub(uint32_t *lut, uint32_t root, const Canon16 &c, const uint16_t *sorted, uint32_t &sub, int lane) {
  const uint32_t i0 = c.base[K + 1], n = c.nsyms - i0;              // (count[] is zero beyond maxlen: base[K+1] is right even then)
  for (int pass = 0; pass < 2; pass++) {
    for (uint32_t g = 0; g < n; g += 32u) {
      const uint32_t i = i0 + g + lane;
      const bool valid = g + lane < n;
      // the code of sorted symbol i and of its right neighbour
      uint32_t L = 0, code = 0, P = 0xffffffffu, Pn = 0xfffffffeu;
      if (valid) {
        for (uint32_t l = K + 1; l <= 15u; l++) {
          if (i - c.base[l] < c.count[l]) { L = l; code = c.first[l] + (i - c.base[l]); P = code >> (l - K); }
          if (i + 1u - c.base[l] < c.count[l]) Pn = (uint32_t)(c.first[l] + (i + 1u - c.base[l])) >> (l - K);
        }
      }
      const uint32_t ri = root + (__brev(P) >> (32 - K));           // the prefix as the stream presents it: first bit lowest
      if (pass == 0) {
        const bool last = valid && P != Pn;
        uint32_t x = last ? 1u << (L - K) : 0u;
        const uint32_t mine = x;
#pragma unroll
        for (int sft = 1; sft < 32; sft <<= 1) {
          const uint32_t u = __shfl_up_sync(TBZ_FULL, x, sft);
          if (lane >= sft) x += u;
        }
        const uint32_t off = sub + x - mine;
        sub += __shfl_sync(TBZ_FULL, x, 31);
        if (last && off + mine <= LUT_N) lut[ri] = K_SUB | (L - K) | (off << 17);
      } else if (valid) {
        const uint32_t link = lut[ri];
        const uint32_t off = link >> 17, sb = link & 31u, r = L - K;
        const uint32_t e = DIST ? d_entry(sorted[i], L) : ll_entry(sorted[i], L);
        for (uint32_t j = __brev(code << (32 - r)) & ((1u << r) - 1u); j < (1u << sb); j += 1u << r) lut[off + j] = e;
      }
    }
    __syncwarp();
    if (sub > LUT_N) return false;
  }
  return true;
}

// What a lane needs only while it looks for the place where it synchronises: kept out of the decode loop's registers
// (the function below is not inlined, so this lives in local memory), as is the code.
struct Rare {
  uint32_t winend, S;                                // the end of the round's window, the sub-chunk size
  uint32_t tgt;                                      // the next place where a synchronisation can happen
  uint32_t j, jstart, c, dof;                        // the lane this one is compared with, its sub-chunk start, its recorded start c (dof: offset of start c - 1)
  uint32_t nx, g_sync;                               // ST_SYNC: synchronised into item g_sync of lane nx
};
// The lane is past its own sub-chunk and about to decode the item that starts at bit p0 >= r.tgt: does an item of the
// lane that owns these bits start here?  Returns the lane's new state (r.tgt: the next place to look at).
__device__ __noinline__ int overrun_event(Rare &r, const WSmem &sm, const uint32_t *gck, uint32_t p0) {
  if (p0 >= r.winend) return ST_END;
  while (p0 >= r.jstart + r.S) { r.j++; r.jstart += r.S; r.c = 1; r.dof = 0; }
  const uint32_t j = r.j, rel = p0 - r.jstart;
  uint32_t c = r.c, dof = r.dof;
  // the first recorded start at or beyond rel: ck = its offset (or "not yet" / "no more"), c its slot
  uint32_t ck = 0u, ckv = 0u;
  if (rel) {
    ck = CK_NONE;
    while (c < CK_DENSE) {
      const uint32_t d = sm.ckd[c][j];
      if (d < 2u) { ck = d ? CK_END : CK_NONE; break; }
      if (dof + d >= rel) { ck = dof + d; break; }
      dof += d; c++;
    }
    if (c >= CK_DENSE)
      while (c < NCK && (ck = (ckv = __ldcg(gck + j * (CK_SPARSE + 1) + (c - CK_DENSE))) & 0xffffu) < rel) c++;
  }
  r.c = c; r.dof = dof;
  if (c < NCK && ck == rel) { r.nx = j; r.g_sync = !rel ? 0u : c < CK_DENSE ? 4u * c : ckv >> 16; return ST_SYNC; }
  if (c < NCK && ck < CK_END) r.tgt = r.jstart + ck;
  else if (c >= NCK || ck == CK_END) r.tgt = r.jstart + r.S;                    // no more recorded starts in that sub-chunk
  else r.tgt = p0;                                                              // its owner is not there yet: look again
  return ST_RUN;
}

// ------------------------------------------------------------------------------------------------
// Every block of a member from bit `pos` on, one warp.  Returns true when the token stream is complete (rec filled in),
// false when the member goes to the sequential kernel.  Every return value is warp-uniform.
// ------------------------------------------------------------------------------------------------
__device__ inline bool decode_blocks(const In &in, uint32_t pos, P1Rec &rec, WSmem &sm, uint32_t *__restrict__ gck,
                                     uint32_t *__restrict__ slabs, uint32_t nslabs, uint32_t *slab_counter, int lane) {
  uint32_t first_slab = NO_SLAB, prev_slab = NO_SLAB;
  uint32_t prev_block_bits = 0;   // size of the previous block of this member: predicts this one
  uint32_t s_cap = S_MAX;         // longest sub-chunk a lane's token list has room for (learned when a list fills up)
  bool last = false;
  uint32_t *const lut = sm.lut;
  uint8_t *const hlens = reinterpret_cast<uint8_t *>(sm.lut + LUT_SUB);

  while (!last) {
    // ================= block header (deflate.lisp:518-528, :577-669) =================
    if (in.end - pos < 3) { TBZ_HD_WHY("give up"); return false; }
    const uint32_t hdr = peek32(in, pos) & 7;
    pos += 3;
    last = hdr & 1;
    const uint32_t btype = hdr >> 1;
    int hlit, hdist;
    __syncwarp();
    if (btype == 1) {
      hlit = 288; hdist = 32;
      for (int i = lane; i < 320; i += 32) hlens[32 + i] = i < 144 ? 8 : i < 256 ? 9 : i < 280 ? 7 : i < 288 ? 8 : 5;
    } else if (btype == 2) {
      if (in.end - pos < 14) { TBZ_HD_WHY("give up"); return false; }
      const uint32_t v = peek32(in, pos);
      hlit = (v & 31) + 257; hdist = ((v >> 5) & 31) + 1;
      const int ncl = ((v >> 10) & 15) + 4;
      if (in.end - pos < 14u + 3u * ncl) { TBZ_HD_WHY("give up"); return false; }
      if (lane < 19) hlens[lane] = 0;
      __syncwarp();
      if (lane < ncl) hlens[c_clen_order[lane]] = peek32(in, pos + 14 + 3 * lane) & 7;
      __syncwarp();
      int err = warp_canon(hlens, 19, sm.h.c_cl, sm.h.sorted_cl, sm.h.run, lane);
      if (!err && sm.h.c_cl.nsyms == 0) err = TBZ_ERR_INVALID_SYMBOL;
      if (err) { TBZ_HD_WHY("give up"); return false; }
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
        uint32_t q = p >> 5;
        unsigned long long bb = ((unsigned long long)tbzfast::ldw(in, q + 1) << 32 | tbzfast::ldw(in, q)) >> (p & 31u);
        uint32_t bc = 64u - (p & 31u);
        q += 2;
        while (idx < total) {
          if (bc < 32u) { bb |= (unsigned long long)tbzfast::ldw(in, q) << bc; bc += 32u; q++; }
          const uint32_t w = (uint32_t)bb;
          const uint32_t r = sm.h.lut_cl[w & 127];
          if (!r) { err = 1; break; }
          const uint32_t L = r & 15, xb = (r >> 4) & 15, sym = r >> 8;
          p += L + xb;
          bb >>= L + xb; bc -= L + xb;
          if (sym < 16) { hlens[32 + idx] = (uint8_t)sym; idx++; lastlen = (int)sym; continue; }
          const uint32_t extra = (w >> L) & ((1u << xb) - 1);
          int rep, val;
          if (sym == 16) { if (lastlen >= 16) { err = 1; break; } rep = 3 + extra; val = lastlen; }
          else { rep = (sym == 17 ? 3 : 11) + extra; val = 0; lastlen = 0; }
          if (idx + rep > total) { err = 1; break; }
          for (int q2 = 0; q2 < rep; q2++) hlens[32 + idx + q2] = (uint8_t)val;
          idx += rep;
        }
        if (p > in.end) err = 1;                                       // the lengths run past the end of the input
      }
      err = __shfl_sync(TBZ_FULL, err, 0);
      if (err) { TBZ_HD_WHY("give up"); return false; }
      pos = __shfl_sync(TBZ_FULL, p, 0);
    } else if (btype == 0) {
      // ================= stored block (deflate.lisp:532-573): LEN, NLEN, then LEN bytes as they are =================
      // They travel as literal tokens, four bytes each, the lanes taking consecutive slices.
      pos = (pos + 7u) & ~7u;
      if (in.end < pos || in.end - pos < 32u) { TBZ_HD_WHY("give up"); return false; }
      const uint32_t v = peek32(in, pos);
      uint32_t slen = v & 0xffffu;
      if ((slen ^ 0xffffu) != (v >> 16)) { TBZ_HD_WHY("give up"); return false; }                 // deflate.lisp:535
      pos += 32u;
      if (in.end - pos < 8u * slen) { TBZ_HD_WHY("give up"); return false; }                      // the input ends inside the block
      uint32_t bp = pos >> 3;                                          // byte offset of the payload from in.w
      pos += 8u * slen;
      while (slen) {
        uint32_t slab_id = 0;
        if (lane == 0) slab_id = atomicAdd(slab_counter, 1u);
        slab_id = __shfl_sync(TBZ_FULL, slab_id, 0);
        if (slab_id >= nslabs) { TBZ_HD_WHY("give up"); return false; }
        uint32_t *slab = slabs + (size_t)slab_id * SLAB_WORDS;
        SlabHdr *sh = reinterpret_cast<SlabHdr *>(slab);
        const uint32_t nb = slen < (uint32_t)(NL * ITEMCAP * 2) ? slen : (uint32_t)(NL * ITEMCAP * 2);
        const uint32_t per = ((nb + NL - 1) / NL + 1u) & ~1u;          // bytes per lane, even
        const uint32_t lo = per * lane < nb ? per * lane : nb, hi = lo + per < nb ? lo + per : nb;
        const uint32_t cnt = (hi - lo + 1u) / 2u;
        uint32_t *list = slab + SLAB_HDR_WORDS + lane * ITEMCAP;
        for (uint32_t t = 0; t < cnt; t++) {
          const uint32_t a = bp + lo + 2u * t;
          const uint32_t b0 = byte_at(in, a);
          list[t] = lo + 2u * t + 1u < hi ? (I_LIT2 | (byte_at(in, a + 1u) << 8) | b0) : b0;
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
      { TBZ_HD_WHY("give up"); return false; }                        // reserved block type: sequential kernel
    }
    __syncwarp();
    // ================= tables (huffman-tree.lisp:99-218) =================
    if (warp_canon(hlens + 32, hlit, sm.h.c_ll, sm.sorted_ll, sm.h.run, lane)) { TBZ_HD_WHY("give up"); return false; }
    if (warp_canon(hlens + 32 + hlit, hdist, sm.h.c_d, sm.h.sorted_d, sm.h.run, lane)) { TBZ_HD_WHY("give up"); return false; }
    if (sm.h.c_ll.nsyms == 0) { TBZ_HD_WHY("give up"); return false; }
    // a lone code longer than its root table would leave holes in a second-level table: nothing writes such a block
    if ((sm.h.c_ll.nsyms == 1 && sm.h.c_ll.maxlen > KLL) || (sm.h.c_d.nsyms == 1 && sm.h.c_d.maxlen > KD)) { TBZ_HD_WHY("give up"); return false; }
    for (int e = lane; e < (1 << KLL); e += 32) {
      const uint32_t r = canon_lookup(sm.h.c_ll, sm.sorted_ll, (uint32_t)e, 1, KLL);
      lut[e] = r ? ll_entry(r >> 4, r & 15) : E_INVALID;
    }
    for (int e = lane; e < (1 << KD); e += 32) {
      const uint32_t r = canon_lookup(sm.h.c_d, sm.h.sorted_d, (uint32_t)e, 1, KD);
      lut[LUT_D + e] = r ? d_entry(r >> 4, r & 15) : E_INVALID;
    }
    __syncwarp();
    {
      uint32_t sub = LUT_SUB;
      if (sm.h.c_ll.maxlen > KLL && !build_sub<KLL, false>(lut, 0u, sm.h.c_ll, sm.sorted_ll, sub, lane)) { TBZ_HD_WHY("give up"); return false; }
      if (sm.h.c_d.maxlen > KD && !build_sub<KD, true>(lut, LUT_D, sm.h.c_d, sm.h.sorted_d, sub, lane)) { TBZ_HD_WHY("give up"); return false; }
    }
    __syncwarp();

    // ================= rounds over the block's compressed bits =================
    // The end of the block is unknown: assume it is about as long as the previous one (libz cuts
    // blocks by symbol count), else that it runs to the end of the input.
    const uint32_t data_start = pos;
    uint32_t expect = in.end - pos;
    if (prev_block_bits && prev_block_bits + prev_block_bits / 16 < expect) expect = prev_block_bits + prev_block_bits / 16;
    bool block_done = false;
    while (!block_done) {
      // ---- geometry of this round
      const uint32_t P0 = pos;
      uint32_t left = in.end - P0;
      if (expect > pos - data_start && expect - (pos - data_start) < left) left = expect - (pos - data_start);
      const uint32_t nrounds = (left + NL * s_cap - 1) / (NL * s_cap);
      uint32_t S = ((left + nrounds - 1) / nrounds + NL - 1) / NL;
      if (S > s_cap) S = s_cap;
      if (S < S_MIN) S = S_MIN;
      uint32_t winend = P0 + S * NL;
      if (winend > in.end) winend = in.end;
      // ---- a slab for this round's token lists
      uint32_t slab_id = 0;
      if (lane == 0) slab_id = atomicAdd(slab_counter, 1u);
      slab_id = __shfl_sync(TBZ_FULL, slab_id, 0);
      if (slab_id >= nslabs) { TBZ_HD_WHY("give up"); return false; }
      uint32_t *const slab = slabs + (size_t)slab_id * SLAB_WORDS;
      SlabHdr *const sh = reinterpret_cast<SlabHdr *>(slab);
#pragma unroll
      for (uint32_t c = 0; c < CK_DENSE; c++) sm.ckd[c][lane] = 0;
      gck[lane * (CK_SPARSE + 1)] = CK_NONE;             // (a sparse slot is set to "not yet" before the one in front of it is filled)
      uint32_t *const list = slab + SLAB_HDR_WORDS + lane * ITEMCAP;
      __syncwarp();

      const uint32_t *const inw = in.w;
      const uint32_t lastw = in.nwords - 1u;       // (words past the end repeat the last one: a lane that gets there is discarded)
      const uint32_t woff = (uint32_t)(((uintptr_t)inw >> 2) & 3u);
      const uint4 *const inb = reinterpret_cast<const uint4 *>(inw - woff);
      const uint32_t lastc = (lastw + woff) >> 2;
      // ---- the lane's state
      const uint32_t cstart = P0 + S * lane, cend = cstart + S;
      uint32_t q0 = 0, q1 = 0, q2 = 0;  // the items of the current group of four that are not stored yet
      uint32_t nitems = 0;              // items the lane emitted (the iteration count when it stopped)
      uint32_t lastoff = 0;             // offset of the item start recorded last
      bool recording = true;            // the lane is inside its own sub-chunk and records item starts
      uint32_t twi = (cend >> 5) + 3u + woff;   // the reader's word index from which on the next place to synchronise may have been reached
      uint32_t endp = cstart;           // where the lane's list ends
      Rare rare;
      rare.winend = winend; rare.S = S; rare.tgt = cend;
      rare.j = lane + 1; rare.jstart = cend; rare.c = 1; rare.dof = 0;
      rare.nx = 0; rare.g_sync = 0;
      // the bit reader: three stream words, a bit offset below 32 between iterations, word wi the next to move up (word
      // indices count from the 16-byte aligned address at or below in.w; the chunk of word wi and the one behind it are
      // in the lane's slots or on their way there); the reader is at bit 32 (wi - 3 - woff) + bo
      uint32_t w0 = 0, w1 = 0, w2 = 0, wi = 3, bo = 0;
      const uint32_t inq = smem_addr(&sm.inq[0][lane]);
      int st = cstart < winend ? ST_RUN : ST_IDLE;
      if (st == ST_RUN) {
        const uint32_t q = cstart >> 5;
        w0 = __ldg(inw + min(q, lastw)); w1 = __ldg(inw + min(q + 1u, lastw)); w2 = __ldg(inw + min(q + 2u, lastw));
        wi = q + 3u + woff; bo = cstart & 31u;
        // the chunk of word wi, and — unless that word is its first: then the loop asks when it gets there — the next
        cp_async16(inq + ((wi & 4u) << 7), inb + min(wi >> 2, lastc)); cp_async_commit();
        if (wi & 3u) { cp_async16(inq + ((~wi & 4u) << 7), inb + min((wi >> 2) + 1u, lastc)); cp_async_commit(); }
      }
      cp_async_wait<0>();               // (the first words are needed right away; the lists' first stores wait for nothing)
      __syncwarp();

      uint32_t it = 0;                  // iterations of the loop = items of every lane that still runs (uniform)
#ifdef TBZ_HD_TIMING
      const long long t_loop0 = clock64();
#endif
      while (__any_sync(TBZ_FULL, st == ST_RUN)) {
        const bool was = st == ST_RUN;
        // ---- every fourth iteration: the lanes inside their own sub-chunk record where this item starts (every 16th
        // once the dense slots are used up); a full list ends the round for every lane
        if ((it & 3u) == 0u) {
          if (it + 8u >= ITEMCAP) { if (st == ST_RUN) { st = ST_CAP; endp = ((wi - 3u - woff) << 5) + bo; } }
          else if (st == ST_RUN && recording && it) {
            const uint32_t p0 = ((wi - 3u - woff) << 5) + bo;
            const bool dense = it < 4u * CK_DENSE;
            const uint32_t si = (it - 4u * CK_DENSE) >> 4;                        // (sparse index, beyond the dense slots)
            if (p0 >= cend || (!dense && si >= CK_SPARSE)) {
              // out of the own sub-chunk (or of slots): no more item starts from this lane
              recording = false;
              const uint32_t sl = dense ? it >> 2 : CK_DENSE + ((it - 4u * CK_DENSE + 15u) >> 4);
              if (sl < NCK) ck_end(sm, gck, sl, lane);
            } else if (dense || (it & 15u) == 0u) {
              ck_put(sm, gck, dense ? it >> 2 : CK_DENSE + si, lane, p0 - cstart, it, lastoff);
              lastoff = p0 - cstart;
            }
          }
        }
        if (st == ST_RUN) {
          // ---- past the own sub-chunk: does an item of the lane that owns these bits start here?
          if (__builtin_expect(wi >= twi, 0)) {
            const uint32_t p0 = ((wi - 3u - woff) << 5) + bo;
            if (p0 >= rare.tgt) {
              st = overrun_event(rare, sm, gck, p0);
              twi = (rare.tgt >> 5) + 3u + woff;
              if (st != ST_RUN) endp = p0;
            }
          }
        }
        if (st == ST_RUN) {
          // ---- first symbol: lit/len
          const uint32_t x = __funnelshift_r(w0, w1, bo);
          uint32_t e = lut[x & ((1u << KLL) - 1u)];
          if (__builtin_expect((e & K_SPECIAL) != 0u, 0)) {
            if (e & K_SUB) e = lut[(e >> 17) + ((x >> KLL) & ~(0xffffffffu << (e & 31u)))];
            if (e & K_STOP) {
              if (e & K_LEN) st = ST_BAD;
              else { st = ST_EOB; endp = ((wi - 3u - woff) << 5) + bo + (e & 31u); }
            }
          }
          const uint32_t n1 = e_drop(e);
          const bool ism = (e & K_LEN) != 0u;
          // ---- second symbol: the distance, or the next lit/len symbol (kept if it is a literal too)
          const uint32_t o2 = bo + n1;
          const bool up = o2 >= 32u;
          const uint32_t y = __funnelshift_r(up ? w1 : w0, up ? w2 : w1, o2);
          uint32_t e2 = lut[ism ? LUT_D + (y & ((1u << KD) - 1u)) : (y & ((1u << KLL) - 1u))];
          if (__builtin_expect((e2 & K_SUB) != 0u, 0)) e2 = lut[(e2 >> 17) + ((y >> (ism ? KD : KLL)) & ~(0xffffffffu << (e2 & 31u)))];
          const bool two = !ism && (e2 & (K_LEN | K_STOP)) == 0u;
          const uint32_t n2 = (ism || two) ? e_drop(e2) : 0u;
          const uint32_t v1 = e_value(e, x, n1), v2 = e_value(e2, y, n2);
          if (__builtin_expect(ism && (e2 & K_STOP) != 0u && st == ST_RUN, 0)) st = ST_BAD;
          if (st == ST_RUN) {
            // ---- drop the bits; one word moves up when the offset passes 32 (two, rarely)
            bo = o2 + n2;
#pragma unroll
            for (int twice = 0; twice < 2; twice++) {
              if (twice == 0 ? __builtin_expect(bo >= 64u, 0) : bo >= 32u) {
                // entering a chunk: it was asked for a chunk ago and has landed (nothing else is on its way); the other
                // slot — its words were moved up and used — takes the chunk behind this one
                const bool first = (wi & 3u) == 0u;
                if (first) cp_async_wait<0>();
                w0 = w1; w1 = w2; w2 = lds32(inq + ((wi & 4u) << 7) + ((wi & 3u) << 2));
                if (first) { cp_async16(inq + ((~wi & 4u) << 7), inb + min((wi >> 2) + 1u, lastc)); cp_async_commit(); }
                wi++; bo -= 32u;
              }
            }
            // ---- the item; four of them leave with one store
            const uint32_t item = (ism ? I_MATCH : two ? I_LIT2 : 0u) | ((ism || two) ? v2 << 8 : 0u) | v1;
            if ((it & 3u) == 3u) *reinterpret_cast<uint4 *>(list + (it - 3u)) = make_uint4(q0, q1, q2, item);
            q0 = q1; q1 = q2; q2 = item;
          }
        }
        if (__builtin_expect(was && st != ST_RUN, 0)) {
          // the lane stops in front of this item: the items of its last group that are not stored yet
          nitems = it;
          const uint32_t r = it & 3u;
          if (r > 2u) list[it - 3u] = q0;
          if (r > 1u) list[it - 2u] = q1;
          if (r > 0u) list[it - 1u] = q2;
          if (recording) {                                                     // no more item starts from this lane
            const uint32_t sl = it < 4u * CK_DENSE ? (it >> 2) + 1u : CK_DENSE + ((it - 4u * CK_DENSE) >> 4) + 1u;
            if (sl < NCK) ck_end(sm, gck, sl, lane);
          }
        }
        it++;
        __syncwarp();
      }
      cp_async_wait<0>();               // nothing may still be on its way into the slots: the next header builds its tables there
#ifdef TBZ_HD_TIMING
      const long long t_loop1 = clock64();
#endif
      // a lane that decoded past the end of the input has nothing proven to offer (repeated words are read there)
      const uint32_t nx = rare.nx, g_sync = rare.g_sync;
      if (st != ST_IDLE && endp > in.end) st = ST_BAD;
      __syncwarp();
      // ---- lanes reachable from lane 0 through "synchronised into" edges are proven
      uint32_t my_g = 0;
      bool proven = false;
      int term_st, cur = 0;
      uint32_t term_pos;
      {
        for (;;) {
          if (lane == cur) proven = true;
          const int st_c = __shfl_sync(TBZ_FULL, st, cur);
          if (st_c != ST_SYNC) { term_st = st_c; term_pos = __shfl_sync(TBZ_FULL, endp, cur); break; }
          const uint32_t nx_c = __shfl_sync(TBZ_FULL, nx, cur);
          const uint32_t g_c = __shfl_sync(TBZ_FULL, g_sync, cur);
          if ((uint32_t)lane == nx_c) my_g = g_c;
          cur = (int)nx_c;
        }
      }
      if (term_st == ST_BAD || term_st == ST_IDLE) { TBZ_HD_WHY("give up"); return false; }
      // a lane whose list filled up ends the round early: shorter sub-chunks from here on
      if (term_st == ST_CAP) {
        const uint32_t used = __shfl_sync(TBZ_FULL, endp - cstart, cur);
        s_cap = used - used / 4u;
        if (s_cap > S_MAX) s_cap = S_MAX;
        if (s_cap < S_MIN) s_cap = S_MIN;
      }
      // ---- the round's slab: every lane's proven tokens, where they are
      {
        const uint32_t cnt = proven && nitems > my_g ? nitems - my_g : 0u;
        sh->fc[lane] = cnt ? (my_g | (cnt << 16)) : 0u;
        if (lane == 0) {
          sh->next = NO_SLAB; sh->out_bytes = 0;
          if (prev_slab != NO_SLAB) reinterpret_cast<SlabHdr *>(slabs + (size_t)prev_slab * SLAB_WORDS)->next = slab_id;
        }
        if (first_slab == NO_SLAB) first_slab = slab_id;
        prev_slab = slab_id;
      }
#ifdef TBZ_HD_TIMING
      if (lane == 0 && (P0 & 7u) == 0u) {
        unsigned smid; asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        printf("T sm %u S %u it %u loop %lld copy %lld start %lld\n", smid, S, it, t_loop1 - t_loop0, clock64() - t_loop1, t_loop0);
      }
#endif
#if defined(TBZ_EMU) && defined(TBZ_EMU_TRACE)
      fprintf(stderr, "[hd]   lane %d st %d cstart %u endp %u k %u nx %u g %u proven %d\n", lane, st, cstart, endp, nitems, nx, g_sync, (int)proven);
      if (lane == 0) fprintf(stderr, "[hd] round P0 %u S %u -> term_st %d pos %u\n", P0, S, term_st, term_pos);
#endif
      // ---- how did the round end?
      if (term_pos <= pos && term_st != ST_EOB) { TBZ_HD_WHY("give up"); return false; }     // no progress (cannot happen; guards the loop)
      pos = term_pos;
      if (term_st == ST_EOB) block_done = true;
      __syncwarp();
    }
    prev_block_bits = pos - data_start;
  }
  if (lane == 0) {
    rec.first_slab = first_slab;
    rec.out_len = 0xffffffffu;             // not counted here: phase two does (and owns the overflow verdict)
    rec.end_pos = pos;
    rec.status = 1u;
  }
  return true;
}

// One member, one warp: wrapper header, then every block.
__device__ inline bool decode_member(const DMember &mem, int fmt, P1Rec &rec, WSmem &sm, uint32_t *__restrict__ gck,
                                     uint32_t *__restrict__ slabs, uint32_t nslabs, uint32_t *slab_counter, int lane) {
  In in;
  uint32_t pos;
  if (!member_start(mem, fmt, in, pos)) { TBZ_HD_WHY("give up"); return false; }
  return decode_blocks(in, pos, rec, sm, gck, slabs, nslabs, slab_counter, lane);
}

}  // namespace tbzhd
