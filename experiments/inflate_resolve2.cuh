// inflate_resolve2.cuh — phase two of the batched fast path, round-2 design: LZ77 resolution of a member's
// token stream (deflate.lisp:244-359 `copy-history`) by ONE WARP per member, no CTA barrier anywhere.
//
// What the round-1 kernel (inflate_copy.cuh: one CTA per member, 6 400-byte windows, dense queues, byte
// pointer jumping between CTA barriers) paid for was its window: 35 % of the match bytes of a window had
// their source inside the window, and every window rebuilt offsets, queues and pointers.  On deflate text
// distances are long (level-6 text of BASELINE config 2: 4 % of the matches reach back less than 128 bytes,
// the median distance is 3.8 KB), so a SMALL unit of work has almost no internal dependency.  Here the unit
// is a step of 32 tokens — one per lane, each `up to four literals + one match` (inflate_decode2.cuh), about
// 280 output bytes:
//   1. the lane's token arrives with one 8-byte load (prefetched a step ahead); a warp scan of the token
//      lengths gives every token its output offset
//   2. literals are stored; a match whose source lies entirely below the step is READY and is copied by its
//      own lane with one straight-line, branch-free sequence — aligned 4-byte loads of the source at
//      immediate offsets, one funnel shift per destination word, 32-bit stores between a <= 3-byte head and
//      tail — the same instructions for every lane whatever the length (<= NFAST bytes) or alignment.
//      Every lane has exactly one match, so no copy slot idles on a literal
//   3. the few matches that reach into the step itself (or overlap their own output: distance < length, the
//      RLE case; or are longer than NFAST; or touch the ring's wrap-around) are then copied in stream order
//      by the whole warp, 32 bytes per pass, the usable distance doubling per pass for overlapping copies
//      (the period trick)
//   4. every 512 finished bytes leave with one 16-byte store per lane; Adler-32 is folded in with dp4a on the
//      way out (order-independent form); gzip's CRC-32 is k_member_crc's job (inflate_crc.cuh)
// The last H bytes of output live in a shared-memory ring per warp (H = 16 KiB: 14 members per SM); a source
// older than that is read back from the member's own output (L2), which by then has been stored.
// Anything irregular — a distance before the start of the output, an output buffer that is too small, a
// trailer that disagrees — sends the member to the sequential kernel, which owns the verdict rules.
#pragma once
#include "tbz_device.cuh"
#include "inflate_decode2.cuh"

namespace tbzr2 {

#if defined(TBZ_EMU) && defined(TBZ_EMU_TRACE)
#define TBZ_R2_WHY(...) do { if (lane == 0) fprintf(stderr, "[r2] " __VA_ARGS__); } while (0)
#else
#define TBZ_R2_WHY(...) do { } while (0)
#endif

using tbzd2::T2_MATCH;
using tbzd2::TOKCAP2;
using tbzfast::NO_SLAB;
using tbzfast::P1Rec;
using tbzfast::SLAB_HDR_WORDS;
using tbzfast::SLAB_WORDS;
using tbzfast::SlabHdr;

#ifndef TBZ_R2_RING
#define TBZ_R2_RING 16384
#endif
#ifndef TBZ_R2_WPC
#define TBZ_R2_WPC 7
#endif
#ifndef TBZ_R2_NFAST
#define TBZ_R2_NFAST 24
#endif
constexpr uint32_t H = TBZ_R2_RING, M = H - 1u;   // ring bytes per warp: absolute output offset p lives at ring[p & M]
constexpr int WPC = TBZ_R2_WPC;                    // warps (members in flight) per CTA
constexpr int NT = WPC * 32;
constexpr uint32_t NFAST = TBZ_R2_NFAST;           // longest match the per-lane straight-line copy takes
constexpr uint32_t NW = NFAST / 4;                 // full destination words of such a match, at most
constexpr uint32_t SBMAX = 32 * (NFAST + 4);       // a step that produces more than this goes token by token
constexpr uint32_t FLUSH = 512;                    // bytes per flush: one 16-byte unit per lane
constexpr uint32_t EDGE = 4 * (NW + 3);            // a token this close to the ring's end takes the ordered path (the fast copy never wraps)
constexpr uint32_t PAD = 16, TAIL = 64;            // shared memory before the first / after the last ring that a fast copy may read (never uses)
static_assert((H & M) == 0 && H >= 4096 && H >= FLUSH + 2 * SBMAX + 1024, "ring margins");
static_assert(NFAST % 4 == 0 && NFAST >= 8 && NFAST <= 32 && EDGE <= TAIL, "straight-line copy length");
constexpr size_t SMEM_BYTES = PAD + (size_t)WPC * H + TAIL;

// The ring is addressed by 32-bit shared-space addresses through ld.shared / st.shared, not through a generic pointer:
// every access of the straight-line copy is then `register + immediate` with no address arithmetic (with a uint8_t*
// the compiler recomputed a window base + offset for each predicated store: +30 % instructions in the copy).
#ifdef TBZ_EMU
__device__ __forceinline__ uint32_t smem_base() { return 0u; }                      // (the emulator: offsets into the block's buffer)
template <class T> __device__ __forceinline__ T lds(uint32_t a) { return *reinterpret_cast<const T *>(::emu::dyn_smem() + a); }
template <class T> __device__ __forceinline__ void sts(uint32_t a, T v) { *reinterpret_cast<T *>(::emu::dyn_smem() + a) = v; }
__device__ __forceinline__ void sts_low8(uint32_t a, uint32_t v) { sts<uint8_t>(a, (uint8_t)v); }
__device__ __forceinline__ void sts_low16(uint32_t a, uint32_t v) { sts<uint16_t>(a, (uint16_t)v); }
#else
extern __shared__ __align__(16) unsigned char tbz_r2_smem[];
__device__ __forceinline__ uint32_t smem_base() { return (uint32_t)__cvta_generic_to_shared(tbz_r2_smem); }
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
template <> __device__ __forceinline__ void sts<uint16_t>(uint32_t a, uint16_t v) { asm volatile("st.shared.u16 [%0], %1;" ::"r"(a), "h"(v) : "memory"); }
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

// Where a lane's source words come from, decided a step ahead (prepare) so that the words of a source older than the
// ring — an L2 round trip — travel while the previous step is still being copied.
struct Src {
  uint32_t E, S[NW + 2];                  // E = the word before slot 0 (the head may start there); S[j], S[j+1] feed word slot j
};

// The geometry of a ready match: hb head bytes up to the first aligned destination word, then full words, then a tail.
__device__ __forceinline__ uint32_t copy_s0(uint32_t dst, uint32_t src) {    // aligned source offset of word slot 0
  return (src & ~3u) + (((src & 3u) + ((0u - dst) & 3u)) & 4u);
}

// The source words of a ready match whose source is older than the ring, read from the member's output (which has them
// by now).  Only words that hold a needed byte are touched.
template <bool AL>
__device__ __forceinline__ void load_far(Src &f, const uint8_t *out, uint32_t dst, uint32_t src, uint32_t n) {
  const uint32_t s0 = copy_s0(dst, src);
  const uint32_t need = src + n - s0;                                  // words at or beyond this byte offset hold nothing needed
  const bool pre = s0 > src;                                           // the head starts in the word before slot 0
  if (AL) {
    const uint32_t *gp = reinterpret_cast<const uint32_t *>(out + s0);
    if (pre) f.E = __ldcg(gp - 1);
#pragma unroll
    for (uint32_t i = 0; i < NW + 2; i++) if (4u * i < need) f.S[i] = __ldcg(gp + i);
  } else {                                                             // `out` is not aligned: byte by byte
    f.E = 0;
    if (pre)
      for (int b = 0; b < 4; b++) f.E |= (uint32_t)__ldcg(out + s0 - 4 + b) << (8 * b);
#pragma unroll
    for (uint32_t i = 0; i < NW + 2; i++) {
      uint32_t v = 0;
      for (uint32_t b = 0; b < 4; b++) if (4u * i + b < need) v |= (uint32_t)__ldcg(out + s0 + 4u * i + b) << (8 * b);
      f.S[i] = v;
    }
  }
}

// One READY match, copied by its own lane: 3 <= n <= NFAST bytes from absolute offset src to dst; the source lies
// entirely below the current step, so it never overlaps the destination, and neither range comes within EDGE bytes of
// the end of the ring.  far: the source is older than the ring; its words are in f already (load_far).
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
      sts<uint8_t>(ring + ((p + lane) & M), (uint8_t)v);
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
      sts<uint8_t>(ring + ((p + done + k) & M), (uint8_t)v);
    }
    __syncwarp();
    done += c;
    back += back;
  }
}

// One whole token by the whole warp, in stream order: its literals, then its match.
__device__ __forceinline__ void token_warp(const WState &w, uint32_t p, uint32_t lo, uint32_t hi, uint32_t ring_lo, int lane) {
  const uint32_t nl = tbzd2::t2_nlit(hi);
  if ((uint32_t)lane < nl) sts<uint8_t>(w.ring + ((p + lane) & M), (uint8_t)(lo >> (8 * lane)));
  __syncwarp();                                  // the match may start with these very bytes
  if (hi & T2_MATCH) copy_warp(w, p + nl, (hi & 255u) + 3u, ((hi >> 8) & 0x7fffu) + 1u, ring_lo, lane);
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

// A step, prepared: the lane's token, where it goes, what kind of work its match is.
struct Prep {
  uint32_t lo, hi, nvalid;                // the token; lanes below nvalid have one
  uint32_t p, nl, n, src;                 // first output byte, literals, match length (0 = none), source offset of the match
  uint32_t base, total;                   // (uniform) where the step starts, how many bytes it produces
  bool ready, far, edge, pend, have;      // match by its own lane / source older than the ring (f holds it) / token near the
  bool fail;                              // ring's end / match for the ordered path / far words already loaded.  fail: uniform
  Src f;
};

// Prepare a step that starts at output offset `base`: offsets by a warp scan, classification, and — for sources older
// than the ring that `out` already holds — the loads of the source words.  flushed: what `out` holds right now.
template <bool AL>
__device__ __forceinline__ void prepare(Prep &q, const WState &w, uint32_t base, uint32_t lo, uint32_t hi, uint32_t nvalid, int lane) {
  q.lo = lo; q.hi = hi; q.nvalid = nvalid; q.base = base;
  const bool v = (uint32_t)lane < nvalid;
  const bool m = v && (hi & T2_MATCH);
  q.nl = v ? tbzd2::t2_nlit(hi) : 0u;
  q.n = m ? (hi & 255u) + 3u : 0u;
  const uint32_t mine = q.nl + q.n;
  uint32_t x = mine;
#pragma unroll
  for (int sft = 1; sft < 32; sft <<= 1) {
    const uint32_t u = __shfl_up_sync(TBZ_FULL, x, sft);
    if (lane >= sft) x += u;
  }
  q.total = __shfl_sync(TBZ_FULL, x, 31);
  q.p = base + x - mine;
  const uint32_t dst = q.p + q.nl;
  const uint32_t d = ((hi >> 8) & 0x7fffu) + 1u;
  q.src = dst - d;
  const uint32_t end = base + q.total;
  // output overflow, or a distance that reaches before the start of the output (deflate.lisp:343-345): the sequential kernel reports it
  q.fail = (unsigned long long)base + q.total > w.cap || __any_sync(TBZ_FULL, m && d > dst);
  const uint32_t ring_lo = end > H ? end - H : 0u;
  q.far = q.src < ring_lo;
  // a token near the ring's end (its bytes, or its source, would wrap) takes the ordered path as a whole
  q.edge = v && ((q.p & M) > H - EDGE - 4u || (m && !q.far && (q.src & M) > H - EDGE));
  q.ready = m && !q.edge && q.src + q.n <= base && q.n <= NFAST && q.total <= SBMAX;
  q.pend = m && !q.ready;
  q.have = q.ready && q.far && q.src + q.n <= w.flushed;
  if (q.have) load_far<AL>(q.f, w.out, dst, q.src, q.n);
}

// Execute a prepared step.  Returns false when the member must go to the sequential kernel.  Warp-uniform result.
template <bool AL>
__device__ __forceinline__ bool execute(WState &w, Prep &q, bool adler, int lane) {
  if (q.fail) { TBZ_R2_WHY("overflow or distance too far at %u (+%u, cap %llu)\n", q.base, q.total, w.cap); return false; }
  const uint32_t ring = w.ring;
  const uint32_t end = q.base + q.total;
  if (__builtin_expect(q.total <= SBMAX, 1)) {
    const uint32_t ring_lo = end > H ? end - H : 0u;
    const uint32_t dst = q.p + q.nl;
    // literals
    {
      const uint32_t lp = ring + (q.p & M), nle = q.edge ? 0u : q.nl;
      if (nle > 0u) sts_low8(lp, q.lo);
      if (nle > 1u) sts_low8(lp + 1u, q.lo >> 8);
      if (nle > 2u) sts_low8(lp + 2u, q.lo >> 16);
      if (nle > 3u) sts_low8(lp + 3u, q.lo >> 24);
    }
    // a far source the preparation could not load yet (it was not in `out` then; it is now)
    if (__builtin_expect(__any_sync(TBZ_FULL, q.ready && q.far && !q.have), 0)) {
      if (q.ready && q.far && !q.have) load_far<AL>(q.f, w.out, dst, q.src, q.n);
    }
    copy_ready(ring, q.f, q.ready, q.far, dst, q.src, q.n);
    __syncwarp();
    // the rest in stream order, by the whole warp
    uint32_t pm = __ballot_sync(TBZ_FULL, q.edge || q.pend);
    if (pm) {
      const uint32_t em = __ballot_sync(TBZ_FULL, q.edge);
      do {
        const int l = __ffs(pm) - 1;
        pm &= pm - 1u;
        const uint32_t pa = __shfl_sync(TBZ_FULL, q.p, l), ha = __shfl_sync(TBZ_FULL, q.hi, l);
        if ((em >> l) & 1u) token_warp(w, pa, __shfl_sync(TBZ_FULL, q.lo, l), ha, ring_lo, lane);      // (literals too)
        else copy_warp(w, pa + tbzd2::t2_nlit(ha), (ha & 255u) + 3u, ((ha >> 8) & 0x7fffu) + 1u, ring_lo, lane);
      } while (pm);
    }
    w.pos = end;
    if (end - w.flushed >= FLUSH) flush_to<AL>(w, w.flushed + ((end - w.flushed) / FLUSH) * FLUSH, adler, lane);
  } else {
    // a step of long matches (RLE, zeros): token by token, so that the ring never runs more than one token ahead of `out`
    for (int l = 0; l < 32; l++) {
      const uint32_t pa = __shfl_sync(TBZ_FULL, q.p, l), la = __shfl_sync(TBZ_FULL, q.lo, l), ha = __shfl_sync(TBZ_FULL, q.hi, l);
      if ((uint32_t)l >= q.nvalid) break;
      const uint32_t e = pa + tbzd2::t2_outlen(ha);
      token_warp(w, pa, la, ha, e > H ? e - H : 0u, lane);
      if (e - w.flushed >= FLUSH) flush_to<AL>(w, w.flushed + ((e - w.flushed) / FLUSH) * FLUSH, adler, lane);
    }
    w.pos = end;
  }
  return true;
}

// The member's token stream, step by step: slabs in chain order, the 32 lists of a slab in lane order, 32 tokens at a time.
struct Cursor {
  const uint32_t *slabs, *slab;
  const uint2 *list;
  uint32_t fc, next_slab, cnt, i0;
  int j;
  __device__ __forceinline__ void open(const uint32_t *slabs_, uint32_t first) {
    slabs = slabs_; slab = nullptr; list = nullptr; cnt = 0; i0 = 0; j = 32; next_slab = first; fc = 0;
  }
  // the next list that has tokens, in the next slab when this one is done; false at the end of the stream.  Uniform.
  __device__ __noinline__ bool next_list(int lane) {
    for (;;) {
      if (j < 31) {
        j++;
        const uint32_t f = __shfl_sync(TBZ_FULL, fc, j);
        cnt = f >> 16; i0 = 0;
        list = reinterpret_cast<const uint2 *>(slab + SLAB_HDR_WORDS) + (uint32_t)j * TOKCAP2 + (f & 0xffffu);
        if (cnt) return true;
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
  // the next step: the lane's token (zero beyond the step's nvalid <= 32 tokens); false at the end of the stream.  Uniform.
  __device__ __forceinline__ bool next(uint2 &t, uint32_t &nvalid, int lane) {
    if (i0 >= cnt && !next_list(lane)) { nvalid = 0; t = make_uint2(0u, 0u); return false; }
    nvalid = cnt - i0 < 32u ? cnt - i0 : 32u;
    t = (uint32_t)lane < nvalid ? __ldg(list + i0 + lane) : make_uint2(0u, 0u);
    i0 += 32u;
    return true;
  }
};

// Every step of the stream, software-pipelined: while step k is copied, step k + 1 is prepared (its far sources are
// on their way) and the tokens of step k + 2 are loaded.
template <bool AL>
__device__ inline bool resolve_stream(WState &w, const P1Rec &rec, const uint32_t *__restrict__ slabs, bool adler, int lane) {
  Cursor cur;
  cur.open(slabs, rec.first_slab);
  uint2 t1, t2;
  uint32_t nv1 = 0, nv2 = 0;
  Prep q, qn;
  bool have0 = cur.next(t1, nv1, lane);
  if (have0) prepare<AL>(q, w, 0u, t1.x, t1.y, nv1, lane);
  bool have1 = have0 && cur.next(t1, nv1, lane);
  while (have0) {
    const bool have2 = have1 && cur.next(t2, nv2, lane);                   // the tokens two steps ahead travel
    if (have1) prepare<AL>(qn, w, q.base + q.total, t1.x, t1.y, nv1, lane);  // the next step's far sources travel
    if (!execute<AL>(w, q, adler, lane)) return false;
    q = qn; have0 = have1; have1 = have2; t1 = t2; nv1 = nv2;
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
__device__ inline bool resolve_member(const DMember &mem, int fmt, const P1Rec &rec, const uint32_t *__restrict__ slabs,
                                      tbz_result &res, uint32_t ring, int lane) {
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
