// inflate_resolve2.cuh — phase two of the batched fast path, round-2 design: LZ77 resolution of a member's
// token stream (deflate.lisp:244-359 `copy-history`) by ONE WARP per member, no CTA barrier anywhere.
//
// What the round-1 kernel (inflate_copy.cuh: one CTA per member, 6 400-byte windows, dense queues, byte
// pointer jumping between CTA barriers) paid for was its window: 35 % of the match bytes of a window had
// their source inside the window, and every window rebuilt offsets, queues and pointers.  On deflate text
// distances are long (level-6 text of BASELINE config 2: 4 % of the matches reach back less than 128 bytes,
// the median distance is 3.8 KB), so a SMALL unit of work has almost no internal dependency.  Here the unit
// is a step of 64 tokens (two per lane, about 270 output bytes):
//   1. the lane's two tokens arrive with one 8-byte load (prefetched a step ahead); a warp scan of their
//      lengths gives every token its output offset
//   2. literals are stored; a match whose source lies entirely below the step is READY and is copied by
//      its own lane with one straight-line sequence — aligned 4-byte loads of the source, one funnel shift
//      per destination word, 32-bit stores between a <= 3-byte head and tail — the same instructions for
//      every lane whatever the length (<= NFAST bytes) or alignment
//   3. the few matches that reach into the step itself (or overlap their own output: distance < length, the
//      RLE case; or are longer than NFAST) are then copied in stream order by the whole warp, 32 bytes per
//      pass, the usable distance doubling per pass for overlapping copies (the period trick)
//   4. every 512 finished bytes leave with one 16-byte store per lane; Adler-32 is folded in with dp4a on the
//      way out (order-independent form); gzip's CRC-32 is k_member_crc's job (inflate_crc.cuh)
// The last H bytes of output live in a shared-memory ring per warp (H = 16 KiB: 14 members per SM); a source
// older than that is read back from the member's own output (L2), which by then has been stored.
// Anything irregular — a distance before the start of the output, an output buffer that is too small, a
// trailer that disagrees — sends the member to the sequential kernel, which owns the verdict rules.
#pragma once
#include "tbz_device.cuh"
#include "inflate_decode.cuh"

namespace tbzr2 {

#if defined(TBZ_EMU) && defined(TBZ_EMU_TRACE)
#define TBZ_R2_WHY(...) do { if (lane == 0) fprintf(stderr, "[r2] " __VA_ARGS__); } while (0)
#else
#define TBZ_R2_WHY(...) do { } while (0)
#endif

using tbzfast::NO_SLAB;
using tbzfast::P1Rec;
using tbzfast::SLAB_HDR_WORDS;
using tbzfast::SLAB_WORDS;
using tbzfast::SlabHdr;
using tbzfast::TOKCAP;
using tbzfast::TOK_LIT2;
using tbzfast::TOK_MATCH;

#ifndef TBZ_R2_RING
#define TBZ_R2_RING 16384
#endif
#ifndef TBZ_R2_WPC
#define TBZ_R2_WPC 7
#endif
#ifndef TBZ_R2_NFAST
#define TBZ_R2_NFAST 16
#endif
constexpr uint32_t H = TBZ_R2_RING, M = H - 1u;   // ring bytes per warp: absolute output offset p lives at ring[p & M]
constexpr int WPC = TBZ_R2_WPC;                    // warps (members in flight) per CTA
constexpr int NT = WPC * 32;
constexpr uint32_t NFAST = TBZ_R2_NFAST;           // longest match the per-lane straight-line copy takes
constexpr uint32_t NW = NFAST / 4;                 // full destination words of such a match, at most
constexpr uint32_t SBMAX = 1024;                   // a step that produces more than this goes token by token
constexpr uint32_t FLUSH = 512;                    // bytes per flush: one 16-byte unit per lane
static_assert((H & M) == 0 && H >= 4096 && H >= FLUSH + 2 * SBMAX + 1024, "ring margins");
static_assert(NFAST % 4 == 0 && NFAST >= 8 && NFAST <= 32, "straight-line copy length");
static_assert(TOKCAP % 2 == 0 && SLAB_HDR_WORDS % 2 == 0, "token pairs are loaded with 8-byte loads");

struct WState {
  uint8_t *ring;                          // this warp's ring
  uint8_t *out;                           // the member's output
  unsigned long long cap;                 // bytes the output may take (capped below 2^32)
  uint32_t pos;                           // output bytes produced so far
  uint32_t flushed;                       // output bytes already stored to global memory (a multiple of 16 until the end)
  unsigned long long acc_a, acc_w;        // per lane: sum d and sum i*d over the bytes it flushed (Adler-32)
};

__device__ __forceinline__ uint32_t tok_len(uint32_t t) { return (t & TOK_MATCH) ? (t & 255u) + 3u : 1u + ((t >> 30) & 1u); }
__device__ __forceinline__ uint32_t ring_word(const uint8_t *ring, uint32_t p) { return *reinterpret_cast<const uint32_t *>(ring + (p & M & ~3u)); }

// One READY match, copied by its own lane: n <= NFAST bytes from absolute offset src to dst, the source entirely
// below the current step (so it never overlaps the destination).  `far`: the source is older than the ring and
// is read from `out`.  Straight-line, no data-dependent branch: head bytes up to the first aligned destination
// word, NW word slots, tail bytes; the source words are aligned loads, one funnel shift per word aligns them.
template <bool AL>
__device__ __forceinline__ void copy_ready(const WState &w, bool act, uint32_t dst, uint32_t src, uint32_t n, bool far) {
  uint8_t *const ring = w.ring;
  const uint32_t hb0 = (0u - dst) & 3u;
  const uint32_t hb = hb0 < n ? hb0 : n;            // head bytes
  const uint32_t nw = (n - hb) >> 2;                 // full words
  const uint32_t tb = (n - hb) & 3u;                 // tail bytes
  const uint32_t as = src & 3u;
  const uint32_t q = as + hb;                        // offset of the first full word's source inside the word grid of src
  const uint32_t sh = (q & 3u) * 8u;
  const uint32_t s0 = (src & ~3u) + (q & 4u);        // aligned source offset of word slot 0
  const uint32_t lim = src + n;                      // source words at or beyond this offset hold nothing that is needed
  // source words S[0..NW+1): slot j needs S[j], S[j+1]
  uint32_t S[NW + 2];
#pragma unroll
  for (uint32_t i = 0; i < NW + 2; i++) {
    const uint32_t a = s0 + 4u * i;
    uint32_t v = 0;
    if (act && a < lim) {
      if (!far) v = ring_word(ring, a);
      else if (AL) v = __ldcg(reinterpret_cast<const uint32_t *>(w.out + a));
      else {
#pragma unroll
        for (int b = 0; b < 4; b++) v |= (uint32_t)__ldcg(w.out + a + b) << (8 * b);   // (reads <= 3 bytes beyond lim - 1: inside the output produced so far or its 16-byte slack)
      }
    }
    S[i] = v;
  }
  // head: the first hb bytes of the stream = the bytes at src
  {
    uint32_t e = 0;                                   // the word before slot 0 (only when the head starts in it: q >= 4)
    if (act && (q & 4u)) {
      const uint32_t a = src & ~3u;
      if (!far) e = ring_word(ring, a);
      else if (AL) e = __ldcg(reinterpret_cast<const uint32_t *>(w.out + a));
      else {
#pragma unroll
        for (int b = 0; b < 4; b++) e |= (uint32_t)__ldcg(w.out + a + b) << (8 * b);
      }
    }
    const uint32_t lo = (q & 4u) ? e : S[0], hi = (q & 4u) ? S[0] : S[1];
    const uint32_t hd = __funnelshift_r(lo, hi, as * 8u);            // stream bytes 0..3
    if (act && (hb & 1u)) ring[dst & M] = (uint8_t)hd;
    if (act && (hb & 2u)) *reinterpret_cast<uint16_t *>(ring + ((dst + (hb & 1u)) & M)) = (uint16_t)(hd >> (8u * (hb & 1u)));
  }
  // full words, and the word the tail lies in
  const uint32_t d0 = dst + hb;                                      // aligned
  uint32_t tw = 0;
#pragma unroll
  for (uint32_t j = 0; j <= NW; j++) {
    const uint32_t v = __funnelshift_r(S[j], S[j + 1], sh);
    if (j < NW && act && j < nw) *reinterpret_cast<uint32_t *>(ring + ((d0 + 4u * j) & M)) = v;
    if (j == nw) tw = v;
  }
  {
    const uint32_t ta = d0 + 4u * nw;
    if (act && (tb & 2u)) *reinterpret_cast<uint16_t *>(ring + (ta & M)) = (uint16_t)tw;
    if (act && (tb & 1u)) ring[(ta + (tb & 2u)) & M] = (uint8_t)(tw >> (8u * (tb & 2u)));
  }
}

// One match copied by the whole warp (warp-uniform arguments): n bytes at absolute offset p, distance d.  A pass
// moves up to `back` bytes from `back` bytes earlier; for an overlapping copy (d < n) everything written so far
// repeats with period d, so the usable distance doubles after every pass (deflate.lisp:286-326 special-cases the
// short periods for the same reason).  A source byte older than the ring is read from `out`.
// ring_lo: offsets below it are not in the ring any more (they are in `out`).
__device__ __forceinline__ void copy_warp(const WState &w, uint32_t p, uint32_t n, uint32_t d, uint32_t ring_lo, int lane) {
  uint8_t *const ring = w.ring;
  uint32_t done = 0, back = d;
  while (done < n) {
    const uint32_t c = back < n - done ? back : n - done;
    for (uint32_t k = lane; k < c; k += 32u) {
      const uint32_t a = p + done + k - back;
      const uint32_t v = a < ring_lo ? (uint32_t)__ldcg(w.out + a) : (uint32_t)ring[a & M];
      ring[(p + done + k) & M] = (uint8_t)v;
    }
    __syncwarp();
    done += c;
    back += back;
  }
}

// 16-byte units [w.flushed, upto) leave the ring: stored to `out`, folded into the Adler-32 sums.  upto is a multiple of 16.
template <bool AL>
__device__ __forceinline__ void flush_to(WState &w, uint32_t upto, bool adler, int lane) {
  for (uint32_t u = w.flushed + 16u * lane; u < upto; u += FLUSH) {
    const uint4 v = *reinterpret_cast<const uint4 *>(w.ring + (u & M));
    if (AL) *reinterpret_cast<uint4 *>(w.out + u) = v;
    else {
      const uint32_t w4[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int b = 0; b < 16; b++) w.out[u + b] = (uint8_t)(w4[b >> 2] >> (8 * (b & 3)));
    }
    if (adler) {
      uint32_t sd = __dp4a(v.x, 0x01010101u, 0u); sd = __dp4a(v.y, 0x01010101u, sd);
      sd = __dp4a(v.z, 0x01010101u, sd); sd = __dp4a(v.w, 0x01010101u, sd);
      uint32_t wj = __dp4a(v.x, 0x03020100u, 0u); wj = __dp4a(v.y, 0x07060504u, wj);
      wj = __dp4a(v.z, 0x0b0a0908u, wj); wj = __dp4a(v.w, 0x0f0e0d0cu, wj);
      w.acc_a += sd;
      w.acc_w += (unsigned long long)u * sd + wj;
    }
  }
  if (upto > w.flushed) w.flushed = upto;
  if (__builtin_expect((w.acc_w >> 62) != 0, 0)) w.acc_w %= TBZ_ADLER_MOD;
  __syncwarp();                     // the stores are ordered before any later read of `out` by another lane
}

// One step: the lane's tokens t0, t1 (tokens 2 lane and 2 lane + 1 of the step's nvalid).  Returns false when the
// member must go to the sequential kernel.  Warp-uniform result.
template <bool AL>
__device__ __forceinline__ bool step(WState &w, uint32_t t0, uint32_t t1, uint32_t nvalid, bool adler, int lane) {
  uint8_t *const ring = w.ring;
  const bool v0 = 2u * lane < nvalid, v1 = 2u * lane + 1u < nvalid;
  const uint32_t l0 = v0 ? tok_len(t0) : 0u, l1 = v1 ? tok_len(t1) : 0u;
  const uint32_t mine = l0 + l1;
  uint32_t x = mine;
#pragma unroll
  for (int sft = 1; sft < 32; sft <<= 1) {
    const uint32_t u = __shfl_up_sync(TBZ_FULL, x, sft);
    if (lane >= sft) x += u;
  }
  const uint32_t total = __shfl_sync(TBZ_FULL, x, 31);
  const uint32_t base = w.pos;
  const uint32_t p0 = base + x - mine, p1 = p0 + l0;
  if ((unsigned long long)base + total > w.cap) { TBZ_R2_WHY("overflow base %u total %u cap %llu\n", base, total, w.cap); return false; }                 // output overflow: the sequential kernel reports it
  const bool m0 = v0 && (t0 & TOK_MATCH), m1 = v1 && (t1 & TOK_MATCH);
  const uint32_t d0 = ((t0 >> 8) & 0x7fffu) + 1u, d1 = ((t1 >> 8) & 0x7fffu) + 1u;
  if (__any_sync(TBZ_FULL, (m0 && d0 > p0) || (m1 && d1 > p1))) { TBZ_R2_WHY("distance too far at %u\n", base); return false; }   // deflate.lisp:343-345
  const uint32_t end = base + total;
  uint32_t pm0, pm1;
  if (total <= SBMAX) {
    // literals
    if (v0 && !m0) { ring[p0 & M] = (uint8_t)t0; if (t0 & TOK_LIT2) ring[(p0 + 1u) & M] = (uint8_t)(t0 >> 8); }
    if (v1 && !m1) { ring[p1 & M] = (uint8_t)t1; if (t1 & TOK_LIT2) ring[(p1 + 1u) & M] = (uint8_t)(t1 >> 8); }
    // ready matches, each by its own lane
    const uint32_t s0 = p0 - d0, s1 = p1 - d1;
    const bool r0 = m0 && s0 + l0 <= base && l0 <= NFAST, r1 = m1 && s1 + l1 <= base && l1 <= NFAST;
    const uint32_t ring_lo = end > H ? end - H : 0u;
    copy_ready<AL>(w, r0, p0, s0, l0, s0 < ring_lo);
    copy_ready<AL>(w, r1, p1, s1, l1, s1 < ring_lo);
    __syncwarp();
    pm0 = __ballot_sync(TBZ_FULL, m0 && !r0);
    pm1 = __ballot_sync(TBZ_FULL, m1 && !r1);
    // the rest in stream order, by the whole warp
    uint32_t pm = pm0 | pm1;
    while (pm) {
      const int l = __ffs(pm) - 1;
      pm &= pm - 1u;
      const uint32_t pa = __shfl_sync(TBZ_FULL, p0, l), na = __shfl_sync(TBZ_FULL, l0 | (d0 << 16), l);
      const uint32_t pb = __shfl_sync(TBZ_FULL, p1, l), nb = __shfl_sync(TBZ_FULL, l1 | (d1 << 16), l);
      if ((pm0 >> l) & 1u) copy_warp(w, pa, na & 0xffffu, na >> 16, ring_lo, lane);      // (the ring already holds the whole step)
      if ((pm1 >> l) & 1u) copy_warp(w, pb, nb & 0xffffu, nb >> 16, ring_lo, lane);
    }
    w.pos = end;
    if (end - w.flushed >= FLUSH) flush_to<AL>(w, w.flushed + ((end - w.flushed) / FLUSH) * FLUSH, adler, lane);
  } else {
    // a step of long matches (RLE, zeros): token by token, the ring never holds more than one token of unflushed slack
    for (int l = 0; l < 32; l++) {
#pragma unroll
      for (int s = 0; s < 2; s++) {
        const uint32_t t = __shfl_sync(TBZ_FULL, s ? t1 : t0, l), p = __shfl_sync(TBZ_FULL, s ? p1 : p0, l);
        if (2u * l + s >= nvalid) continue;
        const uint32_t e = p + tok_len(t);
        if (t & TOK_MATCH) copy_warp(w, p, (t & 255u) + 3u, ((t >> 8) & 0x7fffu) + 1u, e > H ? e - H : 0u, lane);
        else {
          if (lane == 0) { ring[p & M] = (uint8_t)t; if (t & TOK_LIT2) ring[(p + 1u) & M] = (uint8_t)(t >> 8); }
          __syncwarp();
        }
        if (e - w.flushed >= FLUSH) flush_to<AL>(w, w.flushed + ((e - w.flushed) / FLUSH) * FLUSH, adler, lane);
      }
    }
    w.pos = end;
  }
  return true;
}

// The member's token stream, step by step: slabs in chain order, the 32 lists of a slab in lane order, 64 tokens at a time.
struct Cursor {
  const uint32_t *slabs, *slab, *list;
  uint32_t fc, next_slab, cnt, i0;
  int j;
  __device__ __forceinline__ void open(const uint32_t *slabs_, uint32_t first, int lane) {
    slabs = slabs_; slab = nullptr; list = nullptr; cnt = 0; i0 = 0; j = 32; next_slab = first; fc = 0;
    (void)lane;
  }
  // the next step: its first token and how many tokens it has (<= 64); false at the end of the stream.  Uniform.
  __device__ __forceinline__ bool next(const uint32_t *&ptr, uint32_t &nvalid, int lane) {
    for (;;) {
      if (i0 < cnt) {
        ptr = list + i0;
        nvalid = cnt - i0 < 64u ? cnt - i0 : 64u;
        i0 += 64u;
        return true;
      }
      if (j < 31) {
        j++;
        const uint32_t f = __shfl_sync(TBZ_FULL, fc, j);
        cnt = f >> 16; i0 = 0;
        list = slab + SLAB_HDR_WORDS + (uint32_t)j * TOKCAP + (f & 0xffffu);
        continue;
      }
      if (next_slab == NO_SLAB) return false;
      slab = slabs + (size_t)next_slab * SLAB_WORDS;
      const SlabHdr *h = reinterpret_cast<const SlabHdr *>(slab);
      next_slab = __ldg(&h->next);
      fc = __ldg(&h->fc[lane]);
      j = -1; cnt = 0; i0 = 0;
    }
  }
};

__device__ __forceinline__ void load_pair(const uint32_t *ptr, uint32_t nvalid, int lane, uint32_t &t0, uint32_t &t1) {
  t0 = 0; t1 = 0;
  const uint32_t i = 2u * lane;
  if (i + 1u < nvalid) {
    const uint2 v = __ldg(reinterpret_cast<const uint2 *>(ptr + i));      // lists, first proven tokens and steps are all even
    t0 = v.x; t1 = v.y;
  } else if (i < nvalid) t0 = __ldg(ptr + i);
}

template <bool AL>
__device__ inline bool resolve_stream(WState &w, const P1Rec &rec, const uint32_t *__restrict__ slabs, bool adler, int lane) {
  Cursor cur;
  cur.open(slabs, rec.first_slab, lane);
  const uint32_t *ptr = nullptr;
  uint32_t nvalid = 0, t0 = 0, t1 = 0;
  bool have = cur.next(ptr, nvalid, lane);
  if (have) load_pair(ptr, nvalid, lane, t0, t1);
  while (have) {
    const uint32_t *nptr = nullptr;
    uint32_t nn = 0, n0 = 0, n1 = 0;
    const bool have_n = cur.next(nptr, nn, lane);
    if (have_n) load_pair(nptr, nn, lane, n0, n1);                        // travels while this step is copied
    if (!step<AL>(w, t0, t1, nvalid, adler, lane)) return false;
    t0 = n0; t1 = n1; nvalid = nn; have = have_n;
  }
  // what is left in the ring: whole units, then the last partial one byte by byte
  flush_to<AL>(w, w.pos & ~15u, adler, lane);
  if (w.flushed + lane < w.pos) {
    const uint32_t p = w.flushed + lane;
    const uint32_t d = w.ring[p & M];
    w.out[p] = (uint8_t)d;
    w.acc_a += d; w.acc_w += (unsigned long long)p * d;
  }
  __syncwarp();
  return true;
}

// One member, one warp.  Returns false when the caller must queue the member for the sequential kernel.
__device__ inline bool resolve_member(const DMember &mem, int fmt, const P1Rec &rec, const uint32_t *__restrict__ slabs,
                                      tbz_result &res, uint8_t *ring, int lane) {
  WState w;
  w.ring = ring; w.out = mem.out;
  w.cap = mem.out_cap < 0xffffffffull ? mem.out_cap : 0xffffffffull;
  w.pos = 0; w.flushed = 0; w.acc_a = 0; w.acc_w = 0;
  const bool adler = fmt == TBZ_ZLIB;
  const bool al = (((uintptr_t)mem.out) & 15u) == 0;
  if (al ? !resolve_stream<true>(w, rec, slabs, adler, lane) : !resolve_stream<false>(w, rec, slabs, adler, lane)) return false;
  const uint32_t pos = w.pos;
  if (rec.out_len != 0xffffffffu && pos != rec.out_len) { TBZ_R2_WHY("out_len %u != %u\n", pos, rec.out_len); return false; }
  // ---- checksum of the whole member (checksums.lisp:18-62, order-independent form)
  uint32_t ck = 0;
  if (adler) {
    unsigned long long a = w.acc_a % TBZ_ADLER_MOD, ww = w.acc_w % TBZ_ADLER_MOD;
#pragma unroll
    for (int sft = 16; sft; sft >>= 1) { a += __shfl_xor_sync(TBZ_FULL, a, sft); ww += __shfl_xor_sync(TBZ_FULL, ww, sft); }
    const unsigned long long N = pos % TBZ_ADLER_MOD, S = a % TBZ_ADLER_MOD;
    const uint32_t s1 = (uint32_t)((1 + S) % TBZ_ADLER_MOD);
    const uint32_t s2 = (uint32_t)((N + N * S + (unsigned long long)TBZ_ADLER_MOD * 4096 - ww % TBZ_ADLER_MOD) % TBZ_ADLER_MOD);
    ck = s1 | (s2 << 16);
  }
  // ---- trailer (zlib.lisp:80-96, gzip.lisp:82-106): any disagreement goes to the sequential kernel
  const uint32_t mis = (uint32_t)((uintptr_t)mem.in & 3);
  const uint8_t *basep = mem.in - mis;
  const uint32_t endb = (mis + (uint32_t)mem.in_len) * 8;
  uint32_t p = (rec.end_pos + 7) & ~7u;
  if (fmt == TBZ_ZLIB) {
    if (endb - p < 32) return false;
    const uint8_t *q = basep + (p >> 3);
    const uint32_t t = ((uint32_t)q[0] << 24) | ((uint32_t)q[1] << 16) | ((uint32_t)q[2] << 8) | q[3];
    if (t != ck) { TBZ_R2_WHY("adler %08x != %08x\n", ck, t); return false; }
    p += 32;
  } else if (fmt == TBZ_GZIP) {
    if (endb - p < 64) return false;              // (k_member_crc compares the CRC-32 and fills in the checksum)
    p += 64;
  }
  if (lane == 0) {
    res.out_len = pos;
    res.in_used = (p - mis * 8 + 7) >> 3;
    res.checksum = ck;
    res.verdict = TBZ_FINISHED;
    res.where = TBZ_AT_BODY;
    res.path = 1;
  }
  return true;
}

}  // namespace tbzr2
