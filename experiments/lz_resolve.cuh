// lz_resolve.cuh — phase two of the batched fast path (round 2): LZ77 resolution of a member's token stream
// (deflate.lisp:244-359 `copy-history`) by ONE WARP per member, no CTA barrier anywhere.
//
// What the round-1 kernel (inflate_copy.cuh: one CTA per member, 6 400-byte windows, dense queues, byte pointer
// jumping between CTA barriers) paid for was its window: 35 % of the match bytes of a window had their source inside
// the window, and every window rebuilt offsets, queues and pointers.  On deflate text distances are long (level-6
// text of BASELINE config 2: 4 % of the matches reach back less than 128 bytes, the median distance is 3.8 KB), so a
// SMALL unit of work has almost no internal dependency.  Here the unit is a step of 32 tokens — one per lane, each
// `up to four literals + one match` (huff_decode.cuh), about 280 output bytes:
//   1. the lane's token arrives with one coalesced 16-byte load (issued two steps ahead; phase one leaves a member's
//      tokens as contiguous blocks, so steps are full) and carries its output offset: phase one, whose lanes walk
//      their tokens one after the other anyway, counted the bytes')
//   2. literals are stored; a match whose source lies entirely below the step is READY and is copied by its own lane
//      with one straight-line, branch-free sequence — aligned 4-byte loads of the source at immediate offsets, one
//      funnel shift per destination word, 32-bit stores between a <= 3-byte head and tail — the same instructions for
//      every lane whatever the length (<= NFAST bytes) or alignment
//   3. the few matches that reach into the step itself (or overlap their own output: distance < length, the RLE case;
//      or are longer than NFAST; or touch the ring's wrap-around) are then copied in stream order by the whole warp,
//      32 bytes per pass, the usable distance doubling per pass for overlapping copies (the period trick)
//   4. every 512 finished bytes leave with one 16-byte store per lane; Adler-32 is folded in with dp4a on the way out
//      (order-independent form); gzip's CRC-32 is k_member_crc's job (inflate_crc.cuh)
// The last H bytes of output live in a shared-memory ring per warp (H = 16 KiB: 14 members per SM, which shared memory
// limits — registers are plentiful at that occupancy).  A source older than that is in the member's own output by
// then: its words are loaded into registers one step ahead, as soon as the scan of that step has placed its tokens,
// and the straight-line copy takes them from there instead of the ring.
// Anything irregular — a distance before the start of the output, an output buffer that is too small, a trailer that
// disagrees — sends the member to the sequential kernel, which owns the verdict rules.
#pragma once
#include "tbz_device.cuh"
#include "huff_decode.cuh"

namespace tbzlz {

#if defined(TBZ_EMU) && defined(TBZ_EMU_TRACE)
#define TBZ_LZ_WHY(...) do { if (lane == 0) fprintf(stderr, "[lz] " __VA_ARGS__); } while (0)
#else
#define TBZ_LZ_WHY(...) do { } while (0)
#endif

using tbzfast::P1Rec;
using tbzhd::NO_BLOCK;
using tbzhd::T_MATCH;

#ifndef TBZ_LZ_RING
#define TBZ_LZ_RING 16384
#endif
#ifndef TBZ_LZ_WPC
#define TBZ_LZ_WPC 7
#endif
#ifndef TBZ_LZ_NFAST
#define TBZ_LZ_NFAST 24
#endif
#ifndef TBZ_LZ_MINBLOCKS
#define TBZ_LZ_MINBLOCKS 2
#endif
constexpr uint32_t H = TBZ_LZ_RING, M = H - 1u;   // ring bytes per warp: absolute output offset p lives at ring[p & M]
constexpr int WPC = TBZ_LZ_WPC;                    // warps (members in flight) per CTA
constexpr int NT = WPC * 32;
constexpr uint32_t NFAST = TBZ_LZ_NFAST;           // longest match the per-lane straight-line copy takes
constexpr uint32_t NW = NFAST / 4;                 // full destination words of such a match, at most
constexpr uint32_t SBMAX = 32 * (NFAST + 4);       // a step that produces more than this goes token by token
constexpr uint32_t FLUSH = 512;                    // bytes per flush: one 16-byte unit per lane
constexpr uint32_t EDGE = 4 * (NW + 3);            // a token this close to the ring's end takes the ordered path (the fast copy never wraps)
constexpr uint32_t PAD = 16, TAIL = 64;            // shared memory before the first / after the last ring that a fast copy may read (never uses)
static_assert((H & M) == 0 && H >= 32 * 262 + SBMAX + FLUSH + 64, "ring margins: a far source is in `out` a step ahead, even behind a step of 258-byte matches");
static_assert(NFAST % 4 == 0 && NFAST >= 8 && NFAST <= 32 && EDGE <= TAIL, "straight-line copy length");
constexpr size_t SMEM_BYTES = PAD + (size_t)WPC * H + TAIL;
static_assert((SMEM_BYTES + 1024) * TBZ_LZ_MINBLOCKS <= 233472, "CTAs per SM");

// The ring is addressed by 32-bit shared-space addresses through ld.shared / st.shared, not through a generic pointer:
// every access of the straight-line copy is then `register + immediate` with no address arithmetic.
#ifdef TBZ_EMU
__device__ __forceinline__ uint32_t smem_base() { return 0u; }                      // (the emulator: offsets into the block's buffer)
template <class T> __device__ __forceinline__ T lds(uint32_t a) { return *reinterpret_cast<const T *>(::emu::dyn_smem() + a); }
template <class T> __device__ __forceinline__ void sts(uint32_t a, T v) { *reinterpret_cast<T *>(::emu::dyn_smem() + a) = v; }
__device__ __forceinline__ void sts_low8(uint32_t a, uint32_t v) { sts<uint8_t>(a, (uint8_t)v); }
__device__ __forceinline__ void sts_low16(uint32_t a, uint32_t v) { sts<uint16_t>(a, (uint16_t)v); }
#else
extern __shared__ __align__(16) unsigned char tbz_lz_smem[];
__device__ __forceinline__ uint32_t smem_base() { return (uint32_t)__cvta_generic_to_shared(tbz_lz_smem); }
template <class T> __device__ __forceinline__ T lds(uint32_t a);
template <> __device__ __forceinline__ uint8_t lds<uint8_t>(uint32_t a) { uint32_t v; asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a) : "memory"); return (uint8_t)v; }
template <> __device__ __forceinline__ uint32_t lds<uint32_t>(uint32_t a) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory"); return v; }
template <> __device__ __forceinline__ uint4 lds<uint4>(uint32_t a) {
  uint4 v; asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a) : "memory"); return v;
}
template <class T> __device__ __forceinline__ void sts(uint32_t a, T v);
template <> __device__ __forceinline__ void sts<uint8_t>(uint32_t a, uint8_t v) { asm volatile("st.shared.u8 [%0], %1;" ::"r"(a), "r"((uint32_t)v) : "memory"); }
__device__ __forceinline__ void sts_low8(uint32_t a, uint32_t v) { asm volatile("st.shared.u8 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }     // the low byte of v
__device__ __forceinline__ void sts_low16(uint32_t a, uint32_t v) { asm volatile("st.shared.b16 [%0], %1;" ::"r"(a), "h"((uint16_t)v) : "memory"); }
template <> __device__ __forceinline__ void sts<uint32_t>(uint32_t a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
#endif

struct WState {
  uint32_t ring;                          // this warp's ring: shared-space address of its first byte
  uint8_t *out;                           // the member's output
  unsigned long long cap;                 // bytes the output may take (capped below 2^32)
  uint32_t pos;                           // output bytes produced so far
  uint32_t flushed;                       // output bytes already stored to global memory (a multiple of 16 until the end)
  unsigned long long acc_a, acc_w;        // per lane: sum d and sum i*d over the bytes it flushed (Adler-32)
};

// The source words of a far match: E = the word before slot 0 (the head may start there); S[j], S[j+1] feed word slot j.
struct Src { uint32_t E, S[NW + 2]; };

// A step: the lane's token and where its bytes go (the scan), and — loaded as soon as that is known — the source words
// of a match older than the ring.
struct Step {
  uint32_t lo, hi;                        // the token (zero for a lane beyond the step's tokens)
  uint32_t p;                             // the token's first output byte
  uint32_t base, total;                   // (uniform) where the step starts, how many bytes it produces
  Src f;
};

// The geometry of a ready match: hb head bytes up to the first aligned destination word, then full words, then a tail.
// Aligned source offset of word slot 0:
__device__ __forceinline__ uint32_t copy_s0(uint32_t dst, uint32_t src) { return (src & ~3u) + (((src & 3u) + ((0u - dst) & 3u)) & 4u); }

// What a step's own lane does with its token: the classification of execute(), computed early (pure arithmetic on the
// placed step) so that it can be issued between the shuffles of the next step's scan.
struct Cls {
  uint32_t nl, n, dst, src, ring_lo;
  bool m, far, edge, ready, bad;
};
__device__ __forceinline__ Cls classify(const Step &q) {
  Cls c;
  const uint32_t end = q.base + q.total;
  c.m = (q.hi & T_MATCH) != 0u;
  c.nl = tbzhd::t_nlit(q.hi); c.n = (q.hi & 255u) + 3u;
  const uint32_t d = ((q.hi >> 8) & 0x7fffu) + 1u;
  c.dst = q.p + c.nl; c.src = c.dst - d;
  c.bad = c.m && d > c.dst;                          // a distance that reaches before the start of the output (deflate.lisp:343-345)
  c.ring_lo = end > H ? end - H : 0u;
  c.far = c.src < c.ring_lo;
  // a token near the ring's end (its bytes, or its source, would wrap) takes the ordered path as a whole
  c.edge = (q.p & M) > H - EDGE - 4u || (c.m && !c.far && (c.src & M) > H - EDGE);
  c.ready = c.m && !c.edge && c.src + c.n <= q.base && c.n <= NFAST;
  return c;
}

// The next step is placed (phase one wrote every token's output offset next to it: no scan) and the current one,
// already placed, is classified.  Then the words of every source of the new step that is older than the ring start
// their way (the member's output has had them for a long time: see the static_assert on H).  cq == nullptr: nothing to
// classify (the very first step).
template <bool AL>
__device__ __forceinline__ void place(Step &q, const WState &w, uint32_t base, uint32_t lo, uint32_t hi, uint32_t off, int lane, const Step *cq, Cls &cc) {
  q.lo = lo; q.hi = hi; q.base = base;
  const bool m = (hi & T_MATCH) != 0u;
  const uint32_t nl = tbzhd::t_nlit(hi);
  const uint32_t n = m ? (hi & 255u) + 3u : 0u;
  // phase one gave every token its output offset: the step ends where its last token ends
  q.p = (lo | hi) ? off : base;
  q.total = __reduce_max_sync(TBZ_FULL, q.p + nl + n) - base;
  if (cq) cc = classify(*cq);
  const uint32_t dst = q.p + nl, d = ((hi >> 8) & 0x7fffu) + 1u;
  const uint32_t end = base + q.total;
  if (m && d <= dst && end > H && dst - d < end - H && n <= NFAST && (unsigned long long)end <= w.cap) {
    const uint32_t src = dst - d, s0 = copy_s0(dst, src);
    if (AL) {
      const uint32_t *gp = reinterpret_cast<const uint32_t *>(w.out + s0);
      q.f.E = s0 ? __ldcg(gp - 1) : 0u;
#pragma unroll
      for (uint32_t i = 0; i < NW + 2; i++) q.f.S[i] = __ldcg(gp + i);      // (beyond the source, still far below the write position)
    } else {                                                             // `out` is not aligned: byte by byte
      q.f.E = 0;
      if (s0)
        for (int b = 0; b < 4; b++) q.f.E |= (uint32_t)__ldcg(w.out + s0 - 4 + b) << (8 * b);
#pragma unroll
      for (uint32_t i = 0; i < NW + 2; i++) {
        uint32_t v = 0;
        for (uint32_t b = 0; b < 4; b++) v |= (uint32_t)__ldcg(w.out + s0 + 4u * i + b) << (8 * b);
        q.f.S[i] = v;
      }
    }
  }
}

// One READY match, copied by its own lane: 3 <= n <= NFAST bytes from absolute offset src to dst; the source lies
// entirely below the current step, so it never overlaps the destination, and neither range comes within EDGE bytes of
// the end of the ring.  far: the source is older than the ring; its words are in f already (place).
// Straight-line: no data-dependent branch, every shared-memory access at an immediate offset.  Lanes without a ready
// match run along (act = false): they load unused words from wherever their garbage points inside the warp's ring.
__device__ __forceinline__ void copy_ready(uint32_t ring, Src &f, bool act, bool far, uint32_t dst, uint32_t src, uint32_t n) {
  const uint32_t hb = (0u - dst) & 3u;               // head bytes up to the first aligned destination word (n >= 3 >= hb)
  const uint32_t as = src & 3u;
  const uint32_t q = as + hb;                        // offset of the first full word's source on the word grid of src
  const uint32_t sh = (q & 3u) * 8u;
  const uint32_t s0 = (src & ~3u) + (q & 4u);        // aligned source offset of word slot 0
  const uint32_t rest = act ? n - hb : 0u;           // bytes in full words and the tail (none for a lane that only runs along)
  if (!far) {
    const uint32_t rp = ring + (s0 & M);
    f.E = lds<uint32_t>(rp - 4u);
#pragma unroll
    for (uint32_t i = 0; i < NW + 2; i++) f.S[i] = lds<uint32_t>(rp + 4u * i);
  }
  // head: stream bytes 0..hb-1 = the bytes at src
  {
    const uint32_t lo = (q & 4u) ? f.E : f.S[0], hi = (q & 4u) ? f.S[0] : f.S[1];
    const uint32_t hd = __funnelshift_r(lo, hi, as * 8u);
    const uint32_t hp = ring + (dst & M);
    if (act && (hb & 1u)) sts_low8(hp, hd);
    if (act && (hb & 2u)) sts_low16(hp + (hb & 1u), hd >> (8u * (hb & 1u)));
  }
  // full words, and the word the tail lies in
  const uint32_t wp = ring + ((dst + hb) & M);
  uint32_t tw = 0;
#pragma unroll
  for (uint32_t j = 0; j <= NW; j++) {
    const uint32_t v = __funnelshift_r(f.S[j], f.S[j + 1], sh);
    if (j < NW && rest >= 4u * (j + 1u)) sts<uint32_t>(wp + 4u * j, v);
    if (j == 0) tw = v;
    else if (rest >= 4u * j) tw = v;                                   // tw = word slot (rest / 4)
  }
  {
    const uint32_t tp = wp + (rest & ~3u);
    if (rest & 2u) sts_low16(tp, tw);
    if (rest & 1u) sts_low8(tp + (rest & 2u), tw >> (8u * (rest & 2u)));
  }
}

// n bytes at absolute offset p copied by the whole warp from distance d (warp-uniform arguments).  A pass moves up
// to `back` bytes from `back` bytes earlier; for an overlapping copy (d < n) everything written so far repeats with
// period d, so the usable distance doubles after every pass (deflate.lisp:286-326 special-cases the short periods for
// the same reason).  ring_lo: offsets below it are not in the ring any more (they are in `out`).
__device__ __forceinline__ void copy_warp(const WState &w, uint32_t p, uint32_t n, uint32_t d, uint32_t ring_lo, int lane) {
  const uint32_t ring = w.ring;
  if (n <= 32u && d >= n) {                                            // the common case: one pass
    if ((uint32_t)lane < n) {
      const uint32_t a = p + lane - d;
      const uint32_t v = a < ring_lo ? (uint32_t)__ldcg(w.out + a) : (uint32_t)lds<uint8_t>(ring + (a & M));
      sts_low8(ring + ((p + lane) & M), v);
    }
    __syncwarp();
    return;
  }
  uint32_t done = 0, back = d;
  while (done < n) {
    const uint32_t c = back < n - done ? back : n - done;
    for (uint32_t k = lane; k < c; k += 32u) {
      const uint32_t a = p + done + k - back;
      const uint32_t v = a < ring_lo ? (uint32_t)__ldcg(w.out + a) : (uint32_t)lds<uint8_t>(ring + (a & M));
      sts_low8(ring + ((p + done + k) & M), v);
    }
    __syncwarp();
    done += c;
    back += back;
  }
}

// One whole token by the whole warp, in stream order: its literals, then its match.
__device__ __forceinline__ void token_warp(const WState &w, uint32_t p, uint32_t lo, uint32_t hi, uint32_t ring_lo, int lane) {
  const uint32_t nl = tbzhd::t_nlit(hi);
  if ((uint32_t)lane < nl) sts_low8(w.ring + ((p + lane) & M), lo >> (8 * lane));
  __syncwarp();                                  // the match may start with these very bytes
  if (hi & T_MATCH) copy_warp(w, p + nl, (hi & 255u) + 3u, ((hi >> 8) & 0x7fffu) + 1u, ring_lo, lane);
}

// 16-byte units [w.flushed, upto) leave the ring: stored to `out`, folded into the Adler-32 sums.  upto is a multiple of 16.
template <bool AL>
__device__ __forceinline__ void flush_to(WState &w, uint32_t upto, bool adler, int lane) {
  for (uint32_t u = w.flushed + 16u * lane; u < upto; u += FLUSH) {
    const uint4 v = lds<uint4>(w.ring + (u & M));
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

// Execute a placed and classified step.  Returns false when the member must go to the sequential kernel.  Uniform.
template <bool AL>
__device__ __forceinline__ bool execute(WState &w, Step &q, const Cls &c, bool adler, int lane) {
  const uint32_t ring = w.ring;
  const uint32_t end = q.base + q.total;
  // output overflow, or a distance that reaches too far back: the sequential kernel reports it
  if (__any_sync(TBZ_FULL, c.bad) || (unsigned long long)end > w.cap) { TBZ_LZ_WHY("overflow or distance too far at %u (+%u, cap %llu)\n", q.base, q.total, w.cap); return false; }
  if (__builtin_expect(q.total <= SBMAX, 1)) {
    // literals
    {
      const uint32_t lp = ring + (q.p & M), nle = c.edge ? 0u : c.nl;
      if (nle > 0u) sts_low8(lp, q.lo);
      if (nle > 1u) sts_low8(lp + 1u, q.lo >> 8);
      if (nle > 2u) sts_low8(lp + 2u, q.lo >> 16);
      if (nle > 3u) sts_low8(lp + 3u, q.lo >> 24);
    }
    copy_ready(ring, q.f, c.ready, c.far, c.dst, c.src, c.n);
    __syncwarp();
    // the rest in stream order, by the whole warp
    uint32_t pm = __ballot_sync(TBZ_FULL, (c.m && !c.ready) || (c.edge && c.nl));
    if (pm) {
      const uint32_t em = __ballot_sync(TBZ_FULL, c.edge);
      do {
        const int l = __ffs(pm) - 1;
        pm &= pm - 1u;
        const uint32_t pa = __shfl_sync(TBZ_FULL, q.p, l), ha = __shfl_sync(TBZ_FULL, q.hi, l);
        if ((em >> l) & 1u) token_warp(w, pa, __shfl_sync(TBZ_FULL, q.lo, l), ha, c.ring_lo, lane);      // (literals too)
        else copy_warp(w, pa + tbzhd::t_nlit(ha), (ha & 255u) + 3u, ((ha >> 8) & 0x7fffu) + 1u, c.ring_lo, lane);
      } while (pm);
    }
    w.pos = end;
    if (end - w.flushed >= FLUSH) flush_to<AL>(w, w.flushed + ((end - w.flushed) / FLUSH) * FLUSH, adler, lane);
  } else {
    // a step of long matches (RLE, zeros): token by token, so that the ring never runs more than one token ahead of `out`
    for (int l = 0; l < 32; l++) {
      const uint32_t pa = __shfl_sync(TBZ_FULL, q.p, l), la = __shfl_sync(TBZ_FULL, q.lo, l), ha = __shfl_sync(TBZ_FULL, q.hi, l);
      const uint32_t e = pa + tbzhd::t_outlen(ha);
      if (e == pa) continue;                       // (no token in this lane)
      token_warp(w, pa, la, ha, e > H ? e - H : 0u, lane);
      if (e - w.flushed >= FLUSH) flush_to<AL>(w, w.flushed + ((e - w.flushed) / FLUSH) * FLUSH, adler, lane);
    }
    w.pos = end;
  }
  return true;
}

// The member's token stream: the blocks of its chain in order, 32 tokens at a time (phase one: huff_decode.cuh).
struct Cursor {
  const uint4 *heap;
  const uint4 *tok;
  uint32_t left, next;
  __device__ __forceinline__ void open(const uint4 *heap_, uint32_t first) { heap = heap_; tok = nullptr; left = 0; next = first; }
  // the next step: the lane's token (zero beyond the step's tokens); false at the end of the stream.  Uniform.
  __device__ __forceinline__ bool step(uint4 &t) {
    while (left == 0) {
      if (next == NO_BLOCK) { t = make_uint4(0u, 0u, 0u, 0u); return false; }
      const uint4 h = __ldg(heap + next);
      tok = heap + next + 1;
      left = h.y; next = h.x;
    }
    const uint32_t nv = left < 32u ? left : 32u;
    const uint32_t lane = threadIdx.x & 31u;
    t = lane < nv ? __ldg(tok + lane) : make_uint4(0u, 0u, 0u, 0u);
    tok += nv; left -= nv;
#ifndef TBZ_EMU
    if (left > 96u + lane) asm volatile("prefetch.global.L2 [%0];" ::"l"(tok + 96 + lane));   // four steps ahead: the heap is cold in L2
#endif
    return true;
  }
};

// Every step of the stream, software-pipelined: while step k is copied, step k + 1 has been placed (its far sources
// are on their way) and the tokens of step k + 2 are loaded.  Two steps per trip, so that no state changes registers.
template <bool AL>
__device__ inline bool resolve_stream(WState &w, const P1Rec &rec, const uint4 *__restrict__ heap, bool adler, int lane) {
  Cursor cur;
  cur.open(heap, rec.first_slab);
  uint4 ta, tb;                                             // the tokens of the next step to place, alternately
  Step a, b;
  Cls c;
  bool more = cur.step(ta);
  if (more) {
    place<AL>(a, w, 0u, ta.x, ta.y, ta.z, lane, nullptr, c);
    more = cur.step(tb);                                    // the tokens of the step after `a`
    for (;;) {
      // `a` is placed; tb holds the tokens of the step after it (if `more`)
      const bool more2 = more && cur.step(ta);
      if (more) place<AL>(b, w, a.base + a.total, tb.x, tb.y, tb.z, lane, &a, c);
      else c = classify(a);
      if (!execute<AL>(w, a, c, adler, lane)) return false;
      if (!more) break;
      more = more2 && cur.step(tb);
      if (more2) place<AL>(a, w, b.base + b.total, ta.x, ta.y, ta.z, lane, &b, c);
      else c = classify(b);
      if (!execute<AL>(w, b, c, adler, lane)) return false;
      if (!more2) break;
    }
  }
  // what is left in the ring: whole units, then the last partial one byte by byte
  flush_to<AL>(w, w.pos & ~15u, adler, lane);
  if (w.flushed + lane < w.pos) {
    const uint32_t p = w.flushed + lane;
    const uint32_t d = lds<uint8_t>(w.ring + (p & M));
    w.out[p] = (uint8_t)d;
    w.acc_a += d; w.acc_w += (unsigned long long)p * d;
  }
  __syncwarp();
  return true;
}

// One member, one warp.  Returns false when the caller must queue the member for the sequential kernel.
__device__ inline bool resolve_member(const DMember &mem, int fmt, const P1Rec &rec, const uint4 *__restrict__ heap,
                                      tbz_result &res, uint32_t ring, int lane) {
  WState w;
  w.ring = ring; w.out = mem.out;
  w.cap = mem.out_cap < 0xffffffffull ? mem.out_cap : 0xffffffffull;
  w.pos = 0; w.flushed = 0; w.acc_a = 0; w.acc_w = 0;
  const bool adler = fmt == TBZ_ZLIB;
  const bool al = (((uintptr_t)mem.out) & 15u) == 0;
  if (al ? !resolve_stream<true>(w, rec, heap, adler, lane) : !resolve_stream<false>(w, rec, heap, adler, lane)) return false;
  const uint32_t pos = w.pos;
  if (rec.out_len != 0xffffffffu && pos != rec.out_len) { TBZ_LZ_WHY("out_len %u != %u\n", pos, rec.out_len); return false; }
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
    if (t != ck) { TBZ_LZ_WHY("adler %08x != %08x\n", ck, t); return false; }
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

}  // namespace tbzlz
