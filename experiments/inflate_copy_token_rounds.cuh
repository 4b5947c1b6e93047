// experiments/inflate_copy_token_rounds.cuh — inflate_copy.cuh (phase two of the batched path) with the variants of the
// third session of round 2, all parity-green under the emulator and on the GPU, all measured SLOWER than the shipped
// kernel (profiles/r2_experiments.txt: r2j, r2k2, r2l, r2z):
//   -DTBZ_CP_ROUNDS=N   N token rounds over the pending queue (a bitmap of pending bytes; a pending match whose source bits
//                       are clear is copied like a ready one) before pointer jumping takes what is left
//   -DTBZ_CP_PJ=0       token rounds only: no byte pointers, 47 KB of shared memory, 4 CTAs per SM
//   -DTBZ_CP_NT=...     threads per CTA (with -DTBZ_CP_TPT=...)
// To build a library with it: copy it over 3bz_b200/csrc/inflate_copy.cuh and pass the macros through tools/variants.py.
// inflate_copy.cuh — phase two of the batched fast path for byte members: LZ77 resolution of a
// token stream (deflate.lisp:244-359 `copy-history`), one thread per token.
//
// One CTA per member.  Shared memory holds the last 32 KiB of output as a ring (final history) and,
// separately, the window being produced, so positions inside a window never wrap.
// A window is the next <= WT tokens (<= WCAP bytes); tokens are never split.  Per window:
//   1. the tokens are loaded (TPT consecutive tokens per thread), a CTA prefix sum over their
//      lengths gives every token its byte offset
//   2. every thread writes its literals and sorts its matches into two dense job queues (warp scan,
//      one shared-memory atomic per warp): READY = the source lies entirely below the window (final
//      history), PENDING = the source reaches into the window
//   3. the ready queue is copied by all threads, one job per thread and step, so the lanes of a
//      warp all run the same straight-line copy: byte moves for the <= 3 bytes up to the first
//      aligned destination word and after the last one, in between one aligned word load per 4
//      source bytes and a funnel shift.  No atomics, no per-byte token search.  The pending queue is
//      expanded to bytes: val[byte] = window offset of its source, and a dense list of those bytes
//   4. the pending bytes (about a third of the bytes on text) are resolved by pointer jumping with
//      uniform control flow: a byte whose source is final copies it; otherwise it adopts the
//      source's pointer (equal bytes), PJ_HOPS hops per level; a level ends at a CTA barrier
//   5. the window is flushed to global memory and appended to the ring with 16-byte stores;
//      Adler-32 is folded in with dp4a (order-independent form).  gzip's CRC-32 and its trailer
//      compare are a kernel of their own (inflate_crc.cuh; -DTBZ_CP_CRC_SEPARATE=0 keeps them here:
//      a table CRC per 16-byte unit with x^(8n) combines)
// The trailer is checked as zlib.lisp:80-96 / gzip.lisp:82-106 do; any disagreement sends the member
// to the sequential kernel, which owns the verdict rules.
#pragma once
#include "tbz_device.cuh"
#include "inflate_decode.cuh"

#if defined(TBZ_EMU) && defined(TBZ_CP_STATS)
#include <atomic>
#include <cstdio>
namespace tbzcp_stats {
struct S { std::atomic<unsigned long long> tok[8], byt[8]; ~S() { for (int i = 0; i < 8; i++) if (tok[i]) fprintf(stderr, "[cp stats] slot %d: %llu tokens %llu bytes\n", i, (unsigned long long)tok[i], (unsigned long long)byt[i]); } };
static S g;
}
#define TBZ_CP_STAT(slot, n) do { tbzcp_stats::g.tok[(slot) & 7]++; tbzcp_stats::g.byt[(slot) & 7] += (n); } while (0)
#else
#define TBZ_CP_STAT(slot, n) do { } while (0)
#endif

namespace tbzcp {

using tbzfast::NL;
using tbzfast::NO_SLAB;
using tbzfast::P1Rec;
using tbzfast::SLAB_HDR_WORDS;
using tbzfast::SLAB_WORDS;
using tbzfast::SlabHdr;
using tbzfast::TOKCAP;
using tbzfast::TOK_MATCH;

#ifndef TBZ_CP_NT
#define TBZ_CP_NT 256
#endif
constexpr int NT = TBZ_CP_NT;
constexpr int NWARP = NT / 32;
#ifndef TBZ_CP_TPT
#define TBZ_CP_TPT 4
#endif
#ifndef TBZ_CP_WCAP
#define TBZ_CP_WCAP 6400
#endif
constexpr int TPT = TBZ_CP_TPT;                 // tokens per thread and window
constexpr uint32_t WT = (uint32_t)NT * TPT;     // window tokens
constexpr uint32_t HIST = 32768u;
constexpr uint32_t WCAP = TBZ_CP_WCAP;          // window bytes
constexpr uint32_t HMASK = HIST - 1u;
constexpr uint32_t WB = WCAP;                   // (name shared with the other phase-two variants: sizes the x16 table)
constexpr uint32_t V_FINAL = 0xffffu;
#ifndef TBZ_CP_CRC_UPT
#define TBZ_CP_CRC_UPT 2
#endif
#ifndef TBZ_CP_HOPS
#define TBZ_CP_HOPS 4
#endif
constexpr int PJ_HOPS = TBZ_CP_HOPS;           // pointer hops per level of the pending-byte resolution
#ifndef TBZ_CP_ROUNDS
#define TBZ_CP_ROUNDS 0
#endif
constexpr int TOK_ROUNDS = TBZ_CP_ROUNDS;      // token rounds over the pending queue before what is left goes to pointer jumping
#ifndef TBZ_CP_PJ
#define TBZ_CP_PJ 1
#endif
#ifndef TBZ_CP_MAXROUNDS
#define TBZ_CP_MAXROUNDS 48
#endif
constexpr bool PJ = TBZ_CP_PJ != 0;            // false: token rounds until nothing is pending (no byte pointers: 47 KB of shared memory,
                                               // 4 CTAs per SM); a window that needs more than MAXROUNDS sends the member to the sequential kernel
constexpr int MAXROUNDS = TBZ_CP_MAXROUNDS;
static_assert(PJ || TOK_ROUNDS > 0, "something has to resolve the pending matches");
constexpr int CRC_UPT = TBZ_CP_CRC_UPT;
#ifndef TBZ_CP_CRC_SEPARATE
#define TBZ_CP_CRC_SEPARATE 1
#endif
#ifndef TBZ_CP_GHIST
#define TBZ_CP_GHIST 0
#endif
constexpr bool GHIST = TBZ_CP_GHIST != 0;       // experiment: final history is read back from the output in global memory
                                                // (L2) instead of a 32 KiB shared-memory ring: 43 KB per CTA, 5 CTAs/SM
constexpr bool CRC_SEPARATE = TBZ_CP_CRC_SEPARATE != 0;   // gzip: CRC-32 and trailer compare in k_member_crc (inflate_crc.cuh)         // gzip: consecutive 16-byte units per thread between two GF(2) multiplications
static_assert(WT <= 1024u && WCAP <= 8192u && WCAP % 16u == 0, "queue entry fields");

struct Smem {
  alignas(16) uint8_t ring[GHIST ? 16 : HIST]; // final history: absolute output offset p lives at ring[p & HMASK]
  alignas(16) uint8_t win[16 + WCAP + 16];     // the window: offset r (absolute pos + r) lives at win[(pos & 15) + r], so that
                                               // 16-byte units of the output are 16-byte units here
  alignas(16) uint16_t val[PJ ? WCAP : 8];     // per window byte: V_FINAL, or the window offset of an equal byte
  uint16_t pbytes[PJ ? WCAP : 8];              // the bytes of the pending matches
  uint16_t tokoff[WT < 1024u ? 1026u : WT + 2u]; // window offset of every token; [tokens used] = window size (>= 2 KiB: CRC scratch)
  uint32_t jobs[WT];                           // ready matches from [0] up, pending from [WT - 1] down: token | (distance - 1) << 10
  uint16_t pq[TOK_ROUNDS ? 2 : WT];            // per pending match: where its bytes start in pbytes (token rounds: handed out by nleft)
  uint32_t nready, npk;                        // npk: pending matches | their bytes << 16
  uint32_t pend[WCAP / 32 + 1];                // one bit per window byte: it belongs to a pending match that is not copied yet
  uint32_t nleft;                              // bytes of the pending matches the token rounds left (they go to pointer jumping)
  uint32_t hdr[SLAB_HDR_WORDS];
  uint32_t segstart[NL + 1];                   // flat index of the first token of every list of the current slab
  uint32_t segptr[NL];                         // word offset of that token in the slab
  uint32_t crc_tab[CRC_SEPARATE ? 1 : 256];
  uint32_t x16[CRC_SEPARATE ? 1 : WCAP / 16 + 4];   // x^(8 * 16 k) mod P: shifts a CRC over k 16-byte units
  uint32_t crcw[NWARP];
  uint32_t wscan[NWARP], wscan2[NWARP];
  unsigned long long wsum[NWARP][2];
  uint32_t member;
  int fail;
  uint32_t wsize;
  uint32_t crc;
};

// val[] is read by some threads while others advance it (pointer jumping, step 4): these accesses are CTA-scope
// acquire loads and release stores — a byte is copied only after its source's val reads FINAL, and FINAL is
// released after the byte was written — so the one intended concurrent access of the kernel is a data-race-free
// one in the PTX memory model, not an ordering that `volatile` happens to give.
#ifdef TBZ_EMU
__device__ __forceinline__ uint32_t ld_acquire_u16(const uint16_t *p) { return *reinterpret_cast<const volatile uint16_t *>(p); }
__device__ __forceinline__ void st_release_u16(uint16_t *p, uint32_t v) { *reinterpret_cast<volatile uint16_t *>(p) = (uint16_t)v; }
__device__ __forceinline__ void st_relaxed_u16(uint16_t *p, uint32_t v) { *reinterpret_cast<volatile uint16_t *>(p) = (uint16_t)v; }
#else
__device__ __forceinline__ uint32_t ld_acquire_u16(const uint16_t *p) {
  uint16_t v;
  asm volatile("ld.acquire.cta.shared.u16 %0, [%1];" : "=h"(v) : "r"((uint32_t)__cvta_generic_to_shared(p)) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_u16(uint16_t *p, uint32_t v) {
  asm volatile("st.release.cta.shared.u16 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(p)), "h"((uint16_t)v) : "memory");
}
__device__ __forceinline__ void st_relaxed_u16(uint16_t *p, uint32_t v) {     // (a pointer moves on: it publishes no data)
  asm volatile("st.relaxed.cta.shared.u16 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(p)), "h"((uint16_t)v) : "memory");
}
#endif

__device__ __forceinline__ uint32_t tok_len(uint32_t t) { return (t & TOK_MATCH) ? (t & 255u) + 3u : 1u + ((t >> 30) & 1u); }

// Final history: the ring in shared memory, or (GHIST) the member's own output in global memory, read
// through L2 (__ldcg: the bytes were stored by this CTA before a barrier).
struct Hist {
  const uint8_t *ring; const uint8_t *out;
  __device__ __forceinline__ uint32_t byte(uint32_t p) const { return GHIST ? (uint32_t)__ldcg(out + p) : (uint32_t)ring[p & HMASK]; }
  // the aligned 32-bit word number wi of the history (byte offsets 4 wi .. 4 wi + 3 relative to the word grid of `out` / the ring)
  __device__ __forceinline__ uint32_t word(uint32_t wi, uint32_t lim) const {
    if (GHIST) { return wi < lim ? __ldcg(reinterpret_cast<const uint32_t *>(out - ((uintptr_t)out & 3u)) + wi) : 0u; }
    return reinterpret_cast<const uint32_t *>(ring)[wi & (HMASK >> 2)];
  }
  __device__ __forceinline__ uint32_t grid() const { return GHIST ? (uint32_t)((uintptr_t)out & 3u) : 0u; }   // offset of byte 0 on the word grid
};

// The window itself as a copy source (token rounds over the pending queue): offsets are indices into win[].
struct WinSrc {
  const uint8_t *win;
  __device__ __forceinline__ uint32_t byte(uint32_t p) const { return win[p]; }
  __device__ __forceinline__ uint32_t word(uint32_t wi, uint32_t) const { return reinterpret_cast<const uint32_t *>(win)[wi]; }
  __device__ __forceinline__ uint32_t grid() const { return 0u; }
};

// The pending bitmap: bit r = window byte r belongs to a pending match whose bytes do not exist yet.
template <class F> __device__ __forceinline__ void bits_each(uint32_t a, uint32_t n, F f) {
  const uint32_t end = a + n;
  do {
    const uint32_t b = a & 31u, take = min(32u - b, end - a);
    f(a >> 5, (0xffffffffu >> (32u - take)) << b);
    a += take;
  } while (a < end);
}
__device__ __forceinline__ void bits_set(uint32_t *bm, uint32_t a, uint32_t n) { bits_each(a, n, [&](uint32_t w, uint32_t m) { atomicOr(&bm[w], m); }); }
__device__ __forceinline__ void bits_clear(uint32_t *bm, uint32_t a, uint32_t n) { bits_each(a, n, [&](uint32_t w, uint32_t m) { atomicAnd(&bm[w], ~m); }); }
__device__ __forceinline__ bool bits_any(const uint32_t *bm, uint32_t a, uint32_t n) {
  uint32_t any = 0;
  bits_each(a, n, [&](uint32_t w, uint32_t m) { any |= *reinterpret_cast<const volatile uint32_t *>(&bm[w]) & m; });
  return any != 0u;
}

// One match whose source is final history: n bytes from history offset src to win[dst, dst+n).
// Straight-line for n <= 19: byte moves up to the first aligned destination word and after the last
// one, in between one aligned word load per 4 source bytes and a funnel shift.  lim: (GHIST) number of
// history words that may be read (nothing beyond the bytes produced so far).
template <class H>
__device__ __forceinline__ void copy_hist(uint8_t *win, uint32_t dst, const H &h, uint32_t src, uint32_t n, uint32_t lim) {
  const uint32_t dend = dst + n;
  uint32_t hb = (0u - dst) & 3u;                         // bytes up to the first aligned destination word
  if (hb > n) hb = n;
#pragma unroll
  for (uint32_t b = 0; b < 3; b++)
    if (b < hb) win[dst + b] = (uint8_t)h.byte(src + b);
  uint32_t p = dst + hb;                                 // aligned (or the end)
  const uint32_t sa = src + hb + h.grid();                // offset on the source's word grid
  const uint32_t sh = (sa & 3u) * 8u;
  uint32_t wi = sa >> 2;
  uint32_t lo = h.word(wi, lim);
#pragma unroll
  for (int i = 0; i < 4; i++) {
    if (p + 4u <= dend) {
      const uint32_t hi = h.word(wi + 1u + i, lim);
      *reinterpret_cast<uint32_t *>(win + p) = __funnelshift_r(lo, hi, sh);
      lo = hi;
      p += 4u;
    }
  }
  if (__builtin_expect(p + 4u <= dend, 0)) {             // long matches
    wi += 4u;
    do {
      wi++;
      const uint32_t hi = h.word(wi, lim);
      *reinterpret_cast<uint32_t *>(win + p) = __funnelshift_r(lo, hi, sh);
      lo = hi;
      p += 4u;
    } while (p + 4u <= dend);
  }
#pragma unroll
  for (uint32_t b = 0; b < 3; b++)
    if (p + b < dend) win[p + b] = (uint8_t)h.byte(src + (p + b - dst));
}

// CRC-32 of buf[a, a+m): every thread takes one contiguous slice; slices are merged pairwise with
// x^(8 len) shifts.  Only used when the output pointer is not 16-byte aligned.  All threads must call.
__device__ inline void crc_window(Smem &sm, const uint8_t *buf, uint32_t a, uint32_t m, int tid) {
  const uint32_t seg = (m + NT - 1) / NT;
  uint32_t lo = seg * tid, hi = lo + seg;
  if (lo > m) lo = m;
  if (hi > m) hi = m;
  uint32_t c = 0xffffffffu;
  for (uint32_t p = lo; p < hi; p++) c = (c >> 8) ^ sm.crc_tab[(c ^ buf[a + p]) & 0xff];
  c ^= 0xffffffffu;
  if (lo == hi) c = 0;
  uint32_t len = hi - lo;
  uint32_t shift = crc_x8n(seg);
  uint32_t *s_c = reinterpret_cast<uint32_t *>(sm.tokoff), *s_l = s_c + NT;   // 2 KiB: the token offsets are dead by now
  static_assert(CRC_SEPARATE || sizeof(Smem::tokoff) >= 2 * NT * sizeof(uint32_t), "crc scratch");
  for (int s = 1; s < NT; s <<= 1) {
    s_c[tid] = c; s_l[tid] = len;
    __syncthreads();
    if ((tid & (2 * s - 1)) == 0 && tid + s < NT) {
      const uint32_t oc = s_c[tid + s], ol = s_l[tid + s];
      if (ol) {
        const uint32_t f = (ol == seg * (uint32_t)s) ? shift : crc_x8n(ol);
        c = crc_mulmod(f, c) ^ oc;
        len += ol;
      }
    }
    shift = crc_mulmod(shift, shift);
    __syncthreads();
  }
  if (tid == 0) sm.crc = crc_combine(sm.crc, c, m);
}

// per-member state that lives in registers (uniform unless noted)
struct RState {
  uint32_t pos;                           // output bytes produced so far (window base)
  uint32_t flushed;                       // output bytes already stored to global memory
  uint32_t cap;                           // bytes the output may take (phase one does not count them)
  unsigned long long acc_a, acc_w;        // per thread: Adler sum d, sum i*d over the bytes it flushed
};

// The slab's tokens [f, f + n) in flat order: thread tid gets tokens f + TPT tid + q (0 where there is none).
__device__ __forceinline__ void load_tokens(const uint32_t *__restrict__ slab, uint32_t f, uint32_t n, const Smem &sm, int tid, uint32_t (&tk)[TPT]) {
  // the list that holds the thread's first token: last j with segstart[j] <= g
  const uint32_t g0 = f + tid * TPT;
  uint32_t j = 0;
  if ((uint32_t)tid * TPT < n) {
#pragma unroll
    for (int stp = NL / 2; stp; stp >>= 1)
      if (sm.segstart[j + stp] <= g0) j += stp;
  }
  const uint32_t last = tid * TPT + TPT;                 // one past the thread's last token
  if (last <= n && f + last <= sm.segstart[j + 1]) {     // common: all of them in one list
    const uint32_t *src = slab + sm.segptr[j] + (g0 - sm.segstart[j]);
#pragma unroll
    for (int q = 0; q < TPT; q++) tk[q] = __ldg(src + q);
  } else {
#pragma unroll
    for (int q = 0; q < TPT; q++) {
      const uint32_t idx = tid * TPT + q;
      tk[q] = 0u;
      if (idx < n) {
        const uint32_t g = f + idx;
        while (g >= sm.segstart[j + 1]) j++;
        tk[q] = __ldg(slab + sm.segptr[j] + (g - sm.segstart[j]));
      }
    }
  }
}

// One window: the slab's tokens [f, f + n) in flat order (n <= WT); consumes as many whole tokens
// as fit, returns the number consumed (0xffffffff = the member must go to the sequential kernel).
// tk: in, this window's tokens (load_tokens); out, the next window's — of the `total` tokens of the
// slab — loaded while this one is being copied.  All threads must call; the result is uniform.
__device__ inline uint32_t resolve_window(uint8_t *__restrict__ out, int fmt, const uint32_t *__restrict__ slab, uint32_t f, uint32_t n,
                                          uint32_t total_tokens, uint32_t (&tk)[TPT], RState &rs, Smem &sm, int tid) {
  const int lane = tid & 31, warp = tid >> 5;
  const uint32_t pos = rs.pos;
  const uint32_t mis = pos & 15u;
  uint8_t *const buf = sm.win;
  const uint32_t wb = mis;                              // index of the window's first byte in win[]
  // ---- 1. tokens and their offsets
  uint32_t ln[TPT];
  uint32_t mine = 0;
#pragma unroll
  for (int q = 0; q < TPT; q++) {
    ln[q] = (uint32_t)tid * TPT + q < n ? tok_len(tk[q]) : 0u;
    mine += ln[q];
  }
  uint32_t x = mine;
#pragma unroll
  for (int sft = 1; sft < 32; sft <<= 1) {
    const uint32_t u = __shfl_up_sync(TBZ_FULL, x, sft);
    if (lane >= sft) x += u;
  }
  if (lane == 31) sm.wscan[warp] = x;
  if (tid == 0) { sm.nready = 0; sm.npk = 0; sm.nleft = 0; }
  if (TOK_ROUNDS) for (uint32_t i = tid; i < WCAP / 32u + 1u; i += NT) sm.pend[i] = 0u;
  if (PJ) for (uint32_t i = tid; i < WCAP / 8u; i += NT)        // (the previous window's levels ended at a barrier)
    reinterpret_cast<uint4 *>(sm.val)[i] = make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);
  __syncthreads();
  Hist hist;
  hist.ring = sm.ring; hist.out = out;
  const uint32_t hlim = (pos + (uint32_t)((uintptr_t)out & 3u) + 3u) >> 2;   // history words that hold produced bytes
  if (tid < (int)mis) sm.win[tid] = (uint8_t)hist.byte(pos - mis + tid);   // the unit the previous window ended in
  uint32_t off = 0, total = 0;
#pragma unroll
  for (int w = 0; w < NWARP; w++) { const uint32_t c = sm.wscan[w]; if (w < warp) off += c; total += c; }
  uint32_t st[TPT];
  uint32_t used = 0;                                    // bit q: the token is part of this window
  bool bad = false;
  {
    uint32_t s = off + x - mine;
#pragma unroll
    for (int q = 0; q < TPT; q++) {
      const uint32_t idx = tid * TPT + q;
      st[q] = s;
      if (idx < n && s <= WCAP) {
        sm.tokoff[idx] = (uint16_t)s;                   // (the first token that does not fit: its offset is the window size)
        if (s + ln[q] <= WCAP) {
          used |= 1u << q;
          if ((tk[q] & TOK_MATCH) && ((tk[q] >> 8) & 0x7fffu) + 1u > pos + s) bad = true;   // deflate.lisp:343-345
        } else sm.wsize = s;
      }
      s += ln[q];
    }
  }
  if (bad) sm.fail = 1;
  if (tid == 0 && total <= WCAP) { sm.tokoff[n] = (uint16_t)total; sm.wsize = total; }
  if (total > WCAP) {                                   // count the tokens that fit
    uint32_t c = __popc(used);
#pragma unroll
    for (int sft = 16; sft; sft >>= 1) c += __shfl_xor_sync(TBZ_FULL, c, sft);
    if (lane == 0) sm.wscan2[warp] = c;
  }
  // ---- 2. literals; matches into the ready / pending queues
  uint32_t rmask = 0, pmask = 0, pb = 0;
#pragma unroll
  for (int q = 0; q < TPT; q++) {
    if (used & (1u << q)) {
      const uint32_t t = tk[q], dst = wb + st[q];
      if (!(t & TOK_MATCH)) {
        buf[dst] = (uint8_t)t;
        if (t & tbzfast::TOK_LIT2) buf[dst + 1] = (uint8_t)(t >> 8);
      } else {
        const uint32_t d = ((t >> 8) & 0x7fffu) + 1u;
        const uint32_t reach = d < ln[q] ? d : ln[q];    // source bytes that are not the token's own output
        if (d >= st[q] + reach) rmask |= 1u << q;        // entirely below the window
        else { pmask |= 1u << q; pb += ln[q]; if (TOK_ROUNDS) bits_set(sm.pend, st[q], ln[q]); }
      }
    }
  }
  {
    // one warp scan for three counts: ready matches | pending matches << 8 | pending bytes << 16
    const uint32_t c = __popc(rmask) | (__popc(pmask) << 8) | (pb << 16);
    uint32_t incl = c;
#pragma unroll
    for (int sft = 1; sft < 32; sft <<= 1) {
      const uint32_t u = __shfl_up_sync(TBZ_FULL, incl, sft);
      if (lane >= sft) incl += u;
    }
    uint32_t br = 0, bpk = 0;
    if (lane == 31) {
      if (incl & 0xffu) br = atomicAdd(&sm.nready, incl & 0xffu);
      if (incl >> 8) bpk = atomicAdd(&sm.npk, ((incl >> 8) & 0xffu) | ((incl >> 16) << 16));   // matches | bytes << 16: one order for both
    }
    br = __shfl_sync(TBZ_FULL, br, 31); bpk = __shfl_sync(TBZ_FULL, bpk, 31);
    const uint32_t ex = incl - c;
    uint32_t ri = br + (ex & 0xffu), pi = (bpk & 0xffffu) + ((ex >> 8) & 0xffu), qb = (bpk >> 16) + (ex >> 16);
#pragma unroll
    for (int q = 0; q < TPT; q++) {
      const uint32_t job = (tid * TPT + q) | (((tk[q] >> 8) & 0x7fffu) << 10);
      if (rmask & (1u << q)) sm.jobs[ri++] = job;
      if (pmask & (1u << q)) { sm.jobs[WT - 1u - pi] = job; if (!TOK_ROUNDS) sm.pq[pi] = (uint16_t)qb; pi++; qb += ln[q]; }
    }
  }
  __syncthreads();
  uint32_t nused = n;
  if (total > WCAP) {
    nused = 0;
#pragma unroll
    for (int w = 0; w < NWARP; w++) nused += sm.wscan2[w];
  }
  if (sm.fail) return 0xffffffffu;
  const uint32_t wsize = sm.wsize;
  if ((unsigned long long)rs.pos + wsize > rs.cap) return 0xffffffffu;     // output overflow: the sequential kernel reports it
  // ---- 3. the ready queue: one job per thread and step; the pending queue becomes bytes with pointers
  {                                                     // the next window's tokens travel while this one is copied
    const uint32_t fn = f + nused;
    const uint32_t nn = total_tokens - fn < WT ? total_tokens - fn : WT;
    load_tokens(slab, fn, nn, sm, tid, tk);
  }
  {
    const uint32_t nr = sm.nready, np = sm.npk & 0xffffu;
    for (uint32_t j = tid; j < nr; j += NT) {
      const uint32_t job = sm.jobs[j], idx = job & 1023u;
      const uint32_t o = sm.tokoff[idx], d = (job >> 10) + 1u, n_ = sm.tokoff[idx + 1] - o;
      copy_hist(buf, wb + o, hist, pos + o - d, d < n_ ? d : n_, hlim);
      TBZ_CP_STAT(6, n_);
      for (uint32_t k = d; k < n_; k++) buf[wb + o + k] = buf[wb + o + k - d];   // (first token of the window only) its own period
    }
    if (TOK_ROUNDS) {
      // ---- 3b. token rounds: a pending match whose source bytes all exist by now (literals, ready matches, pending
      // matches of an earlier round: no bit of the pending bitmap over its source) is copied like a ready one, from the
      // window; its bits are cleared behind a fence, so a match that finds them clear may read the bytes.  On text a
      // window's dependency chains are two or three tokens deep; what TOK_ROUNDS rounds leave (long chains: runs)
      // becomes bytes with pointers as before
      const WinSrc wsrc{buf};
      __syncthreads();                                  // literals and ready matches are in the window
      for (int round = 0; !PJ || round < TOK_ROUNDS; round++) {
        int left = 0;
        for (uint32_t j = tid; j < np; j += NT) {
          const uint32_t job = sm.jobs[WT - 1u - j];
          if (job >> 31) continue;                      // copied in an earlier round
          const uint32_t idx = job & 1023u, d = ((job >> 10) & 0x7fffu) + 1u;
          const uint32_t s0 = sm.tokoff[idx], n_ = sm.tokoff[idx + 1] - s0;
          const uint32_t m = d < n_ ? d : n_;           // source bytes that are not the token's own output
          if (s0 < d) {                                 // its source begins below the window (rare)
            if (PJ) continue;                           // bytes with pointers
            const uint32_t inwin = s0 + m > d ? s0 + m - d : 0u;
            if (inwin && bits_any(sm.pend, 0u, inwin)) { left = 1; continue; }
            __threadfence_block();
            for (uint32_t k = 0; k < n_; k++) {
              const uint32_t r = s0 + k;
              buf[wb + r] = r < d ? (uint8_t)hist.byte(pos + r - d) : buf[wb + r - d];
            }
          } else {
            if (bits_any(sm.pend, s0 - d, m)) { left = 1; continue; }
            __threadfence_block();
            copy_hist(buf, wb + s0, wsrc, wb + s0 - d, m, 0u);
            for (uint32_t k = d; k < n_; k++) buf[wb + s0 + k] = buf[wb + s0 + k - d];   // its own period
          }
          __threadfence_block();
          bits_clear(sm.pend, s0, n_);
          sm.jobs[WT - 1u - j] = job | 0x80000000u;
          TBZ_CP_STAT(round, n_);
        }
        if (!__syncthreads_or(left)) break;
        if (!PJ && round + 1 >= MAXROUNDS) return 0xffffffffu;   // (uniform) a chain this deep: the sequential kernel copies in order
      }
    }
    if (PJ) for (uint32_t j = tid; j < np; j += NT) {
      const uint32_t job = sm.jobs[WT - 1u - j], idx = job & 1023u, d = ((job >> 10) & 0x7fffu) + 1u;
      if (job >> 31) continue;
      const uint32_t s0 = sm.tokoff[idx], n_ = sm.tokoff[idx + 1] - s0;
      uint16_t *qp = sm.pbytes + (TOK_ROUNDS ? atomicAdd(&sm.nleft, n_) : (uint32_t)sm.pq[j]);
      TBZ_CP_STAT(7, n_);
      for (uint32_t k = 0; k < n_; k++) {
        const uint32_t r = s0 + k;
        qp[k] = (uint16_t)r;
        if (r < d) buf[wb + r] = (uint8_t)hist.byte(pos + r - d);   // the source is below the window: final
        else sm.val[r] = (uint16_t)(r - d);
      }
    }
  }
  // ---- 4. pointer jumping over the pending bytes
  if (PJ) {
    __syncthreads();
    const uint32_t nb = TOK_ROUNDS ? sm.nleft : sm.npk >> 16;
    for (;;) {
      int unresolved = 0;
      for (uint32_t i = tid; i < nb; i += NT) {
        const uint32_t r = sm.pbytes[i];
        uint16_t *vp = &sm.val[r];
        const uint32_t s2 = ld_acquire_u16(vp);
        if (s2 != V_FINAL) {
          uint32_t sp = s2;
          uint32_t vs = ld_acquire_u16(&sm.val[sp]);
#pragma unroll
          for (int hop = 1; hop < PJ_HOPS; hop++)         // several hops per level: fewer levels, fewer barriers
            if (vs != V_FINAL) { sp = vs; vs = ld_acquire_u16(&sm.val[sp]); }
          if (vs == V_FINAL) {
            buf[wb + r] = *reinterpret_cast<volatile uint8_t *>(&buf[wb + sp]);   // (ordered behind the acquire that read FINAL)
            st_release_u16(vp, V_FINAL);                                           // the byte first, then FINAL
          } else { st_relaxed_u16(vp, vs); unresolved = 1; }   // equal bytes: adopt the source's pointer
        }
      }
      if (!__syncthreads_or(unresolved)) break;
    }
  }
  // ---- 5. flush complete 16-byte units, fold them into the checksum
  const bool aligned_out = (((uintptr_t)out) & 15) == 0;
  const bool crc_here = fmt == TBZ_GZIP && !CRC_SEPARATE;
  if (crc_here && !aligned_out) crc_window(sm, buf, wb, wsize, tid);
  if (aligned_out) {
    const uint32_t upto = (pos + wsize) & ~15u;
    const uint8_t *b0 = buf - (pos - mis);               // b0 + absolute offset (16-byte units stay aligned)
    uint32_t myc = 0;                                    // gzip: CRCs of this thread's units, shifted to the end of the flushed range
    if (crc_here) {
      // crc(A || B) = crc(A) * x^(8 |B|) + crc(B) for finalized CRCs: a thread runs the table CRC over
      // CRC_UPT consecutive 16-byte units, then multiplies once by the power for the bytes that follow
      // them (no carry-less multiply on sm_100a: 32 shift-and-xor steps); XOR over all threads
      for (uint32_t p = rs.flushed + 16u * CRC_UPT * tid; p < upto; p += 16u * CRC_UPT * NT) {
        uint32_t c = 0xffffffffu, pe = p;
#pragma unroll
        for (int j = 0; j < CRC_UPT; j++) {
          if (p + 16u * j < upto) {
            const uint4 v = *reinterpret_cast<const uint4 *>(b0 + p + 16u * j);
            *reinterpret_cast<uint4 *>(out + p + 16u * j) = v;
            if (!GHIST) *reinterpret_cast<uint4 *>(&sm.ring[(p + 16u * j) & HMASK]) = v;
            const uint32_t w4[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int q = 0; q < 4; q++)
#pragma unroll
              for (int b8 = 0; b8 < 4; b8++) c = (c >> 8) ^ sm.crc_tab[(c ^ (w4[q] >> (8 * b8))) & 0xff];
            pe = p + 16u * j + 16u;
          }
        }
        myc ^= crc_mulmod(sm.x16[(upto - pe) >> 4], c ^ 0xffffffffu);
      }
    } else {
      for (uint32_t p = rs.flushed + 16u * tid; p < upto; p += 16u * NT) {
        const uint4 v = *reinterpret_cast<const uint4 *>(b0 + p);
        *reinterpret_cast<uint4 *>(out + p) = v;
        if (!GHIST) *reinterpret_cast<uint4 *>(&sm.ring[p & HMASK]) = v;
        if (fmt == TBZ_ZLIB) {
          uint32_t sd = __dp4a(v.x, 0x01010101u, 0u); sd = __dp4a(v.y, 0x01010101u, sd);
          sd = __dp4a(v.z, 0x01010101u, sd); sd = __dp4a(v.w, 0x01010101u, sd);
          uint32_t wj = __dp4a(v.x, 0x03020100u, 0u); wj = __dp4a(v.y, 0x07060504u, wj);
          wj = __dp4a(v.z, 0x0b0a0908u, wj); wj = __dp4a(v.w, 0x0f0e0d0cu, wj);
          rs.acc_a += sd;
          rs.acc_w += (unsigned long long)p * sd + wj;
        }
      }
    }
    if (upto + tid < pos + wsize) {                      // the unit the window ends in: history too
      if (GHIST) out[upto + tid] = b0[upto + tid]; else sm.ring[(upto + tid) & HMASK] = b0[upto + tid];
    }
    if (crc_here) {
#pragma unroll
      for (int sft = 16; sft; sft >>= 1) myc ^= __shfl_xor_sync(TBZ_FULL, myc, sft);
      if (lane == 0) sm.crcw[warp] = myc;
      __syncthreads();
      if (tid == 0 && upto > rs.flushed) {
        uint32_t wc = 0;
#pragma unroll
        for (int w = 0; w < NWARP; w++) wc ^= sm.crcw[w];
        sm.crc = crc_mulmod(sm.x16[(upto - rs.flushed) >> 4], sm.crc) ^ wc;
      }
    }
    if (upto > rs.flushed) rs.flushed = upto;
  } else {
    for (uint32_t p = pos + tid; p < pos + wsize; p += NT) {
      const uint32_t d = buf[wb + p - pos];
      out[p] = (uint8_t)d;
      if (!GHIST) sm.ring[p & HMASK] = (uint8_t)d;
      rs.acc_a += d; rs.acc_w += (unsigned long long)p * d;
    }
    rs.flushed = pos + wsize;
  }
  if (__builtin_expect((rs.acc_w >> 62) != 0, 0)) rs.acc_w %= TBZ_ADLER_MOD;
  rs.pos = pos + wsize;
  // no barrier here: the next window reaches one before it reads the ring or writes anything this flush reads
  return nused;
}

// Every window of one member's token stream.  Returns false when the caller must fall back.
__device__ inline bool resolve_stream(uint8_t *__restrict__ out, int fmt, const P1Rec &rec, const uint32_t *__restrict__ slabs,
                                      RState &rs, Smem &sm, int tid) {
  if (tid == 0) { sm.fail = 0; sm.crc = 0; }
  __syncthreads();
  for (uint32_t s = rec.first_slab; s != NO_SLAB;) {
    const uint32_t *slab = slabs + (size_t)s * SLAB_WORDS;
    if (tid < (int)SLAB_HDR_WORDS) sm.hdr[tid] = slab[tid];
    __syncthreads();
    s = sm.hdr[0];
    if (tid < 32) {                        // flat token order of the slab: exclusive scan of the list sizes
      const uint32_t fc = sm.hdr[4 + tid];
      const uint32_t cnt = fc >> 16;
      uint32_t y = cnt;
#pragma unroll
      for (int sft = 1; sft < 32; sft <<= 1) {
        const uint32_t u = __shfl_up_sync(TBZ_FULL, y, sft);
        if (tid >= sft) y += u;
      }
      sm.segstart[tid] = y - cnt;
      sm.segptr[tid] = SLAB_HDR_WORDS + tid * TOKCAP + (fc & 0xffffu);
      if (tid == 31) sm.segstart[32] = y;
    }
    __syncthreads();
    const uint32_t total = sm.segstart[32];
    uint32_t f = 0;
    uint32_t tk[TPT];
    load_tokens(slab, 0, total < WT ? total : WT, sm, tid, tk);
    while (f < total) {
      const uint32_t n = total - f < WT ? total - f : WT;
      const uint32_t used = resolve_window(out, fmt, slab, f, n, total, tk, rs, sm, tid);
      if (used == 0xffffffffu || used == 0) return false;
      f += used;
    }
    __syncthreads();
  }
  if (sm.fail) return false;
  if (rs.flushed + tid < rs.pos) {         // the last partial 16-byte unit
    const uint32_t p = rs.flushed + tid;
    const uint32_t d = GHIST ? (uint32_t)__ldcg(out + p) : (uint32_t)sm.ring[p & HMASK];
    out[p] = (uint8_t)d;
    rs.acc_a += d; rs.acc_w += (unsigned long long)p * d;
  }
  return true;
}

__device__ inline bool resolve_member(const DMember &mem, int fmt, const P1Rec &rec, const uint32_t *__restrict__ slabs,
                                      tbz_result &res, Smem &sm, int tid) {
  const int lane = tid & 31, warp = tid >> 5;
  RState rs;
  rs.pos = 0; rs.flushed = 0; rs.acc_a = 0; rs.acc_w = 0;
  rs.cap = mem.out_cap < 0xffffffffull ? (uint32_t)mem.out_cap : 0xffffffffu;
  if (!resolve_stream(mem.out, fmt, rec, slabs, rs, sm, tid)) return false;
  const uint32_t pos = rs.pos;
  if (rec.out_len != 0xffffffffu && pos != rec.out_len) return false;
  unsigned long long acc_a = rs.acc_a, acc_w = rs.acc_w;
  // ---- checksum of the whole member
  uint32_t ck = 0;
  if (fmt == TBZ_ZLIB) {
    unsigned long long a = acc_a % TBZ_ADLER_MOD, w = acc_w % TBZ_ADLER_MOD;
#pragma unroll
    for (int sft = 16; sft; sft >>= 1) { a += __shfl_xor_sync(TBZ_FULL, a, sft); w += __shfl_xor_sync(TBZ_FULL, w, sft); }
    if (lane == 0) { sm.wsum[warp][0] = a; sm.wsum[warp][1] = w; }
    __syncthreads();
    a = 0; w = 0;
    for (int k = 0; k < NWARP; k++) { a += sm.wsum[k][0]; w += sm.wsum[k][1]; }
    const unsigned long long N = pos % TBZ_ADLER_MOD, S = a % TBZ_ADLER_MOD;
    const uint32_t s1 = (uint32_t)((1 + S) % TBZ_ADLER_MOD);
    const uint32_t s2 = (uint32_t)((N + N * S + (unsigned long long)TBZ_ADLER_MOD * 4096 - w % TBZ_ADLER_MOD) % TBZ_ADLER_MOD);
    ck = s1 | (s2 << 16);
  } else if (fmt == TBZ_GZIP && !CRC_SEPARATE) {
    __syncthreads();
    uint32_t c = sm.crc;
    if (rs.flushed < pos) {                  // the last partial unit (uniform: every thread computes the same value)
      uint32_t t = 0xffffffffu;
      for (uint32_t p = rs.flushed; p < pos; p++) t = (t >> 8) ^ sm.crc_tab[(t ^ (GHIST ? (uint32_t)__ldcg(mem.out + p) : (uint32_t)sm.ring[p & HMASK])) & 0xff];
      c = crc_combine(c, t ^ 0xffffffffu, pos - rs.flushed);
    }
    ck = c;
  }
  // ---- trailer (zlib.lisp:80-96, gzip.lisp:82-106): any disagreement goes to the sequential kernel
  uintptr_t a0 = (uintptr_t)mem.in;
  const uint32_t mis = (uint32_t)(a0 & 3);
  const uint8_t *base = mem.in - mis;
  const uint32_t end = (mis + (uint32_t)mem.in_len) * 8;
  uint32_t p = (rec.end_pos + 7) & ~7u;
  if (fmt == TBZ_ZLIB) {
    if (end - p < 32) return false;
    const uint8_t *q = base + (p >> 3);
    const uint32_t t = ((uint32_t)q[0] << 24) | ((uint32_t)q[1] << 16) | ((uint32_t)q[2] << 8) | q[3];
    if (t != ck) return false;
    p += 32;
  } else if (fmt == TBZ_GZIP) {
    if (end - p < 64) return false;
    const uint8_t *q = base + (p >> 3);
    const uint32_t t = q[0] | ((uint32_t)q[1] << 8) | ((uint32_t)q[2] << 16) | ((uint32_t)q[3] << 24);
    if (!CRC_SEPARATE && t != ck) return false;        // (separate CRC kernel: it compares, and fills in the checksum)
    p += 64;
  }
  if (tid == 0) {
    res.out_len = pos;
    res.in_used = (p - mis * 8 + 7) >> 3;
    res.checksum = ck;
    res.verdict = TBZ_FINISHED;
    res.where = TBZ_AT_BODY;
    res.path = 1;
  }
  return true;
}

}  // namespace tbzcp
