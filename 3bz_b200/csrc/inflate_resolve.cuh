// inflate_resolve.cuh — phase two of the batched fast path: LZ77 resolution of a token stream
// (deflate.lisp:244-359 `copy-history`, restated for a whole CTA).
//
// One CTA per member.  Shared memory holds the 32 KiB deflate history as a ring (hist) and the
// window being produced (win).  A window is one proven token list of phase one (or a piece of it,
// when it would exceed WT tokens or WB bytes):
//   1. token-parallel: the tokens are loaded, a CTA prefix sum over their lengths gives every
//      token its byte offset, and every token sets one bit in a "token starts here" bitmap over
//      the window's bytes; a prefix popcount over the bitmap words makes byte -> token a rank query
//   2. byte-parallel, the same straight-line code for every byte: one aligned output word per
//      thread and step; byte -> token -> literal value, or source = byte - distance (overlapping
//      matches go through their period).  A source below the window is final history and is read
//      at once.  A source inside the window leaves the byte pending: its source pointer is
//      stored in val[] and the byte is queued
//   3. pending bytes are resolved by pointer jumping over val[], dense and balanced because they
//      sit in a queue: a byte whose source is final copies it; otherwise it adopts the source's
//      pointer (equal bytes) and is queued again.  Chains halve every level
//   4. the window is appended to the history ring and flushed to global memory with 16-byte
//      stores; Adler-32 is folded in with dp4a as s1 = 1 + sum d, s2 = N + N sum d - sum i d_i
// CRC-32 (gzip) is a thread-parallel pass per window with x^(8 len) combines.  The trailer is then
// checked as zlib.lisp:80-96 / gzip.lisp:82-106 do; any disagreement sends the member to the
// sequential kernel, which owns the verdict rules.
#pragma once
#include "tbz_device.cuh"
#include "inflate_decode.cuh"

namespace tbzres {

using tbzfast::NL;
using tbzfast::NO_SLAB;
using tbzfast::P1Rec;
using tbzfast::SLAB_HDR_WORDS;
using tbzfast::SLAB_WORDS;
using tbzfast::SlabHdr;
using tbzfast::TOKCAP;
using tbzfast::TOK_MATCH;

// 512 threads, two tokens each, 4 096-element windows: the kernel's only instance resolves 16-bit symbols with a 64 KB
// ring (110 KB: two CTAs per SM), and with 256 threads per CTA those were 16 warps per SM at IPC 1.7 with a quarter of the
// stall samples at barriers; 32 warps: 7.45 -> 6.10 ms per GiB (384 threads x 3 tokens: 6.94; gpurun_out/r2rv.log)
#ifndef TBZ_RES_NT
#define TBZ_RES_NT 512
#endif
constexpr int NT = TBZ_RES_NT;
constexpr int NWARP = NT / 32;
constexpr uint32_t HIST = 32768u, HMASK = HIST - 1u;
constexpr bool CRC_SEPARATE = false;    // CRC-32 inside the kernel
#ifndef TBZ_RES_TPT
#define TBZ_RES_TPT 2
#endif
constexpr int TPT = TBZ_RES_TPT;         // tokens per thread and window
#ifndef TBZ_RES_WB
#define TBZ_RES_WB 4096
#endif
constexpr uint32_t WB = TBZ_RES_WB;      // window bytes (including the <= 3 bytes of alignment lead-in)
static_assert(WB % 1024u == 0 && WB <= 8192u, "rank directory: 32 lanes x WB / 1024 bitmap words");
constexpr uint32_t WT = TPT * NT;        // window tokens
constexpr uint32_t V_FINAL = 0xffffu;

// T = uint8_t: bytes.  T = uint16_t: symbols of the split decode of one large member — a byte, or
// SYM_MARK | i for "the byte i positions into the 32 KiB that precede this chunk" (unknown until the
// previous chunk is final).  Everything below moves T around without looking inside.
constexpr uint32_t SYM_MARK = 0x8000u;
template <typename T> struct Vec4;
template <> struct Vec4<uint8_t> { typedef uint32_t type; };
template <> struct Vec4<uint16_t> { typedef uint2 type; };

template <typename T>
struct SmemT {
  alignas(16) T hist[HIST];              // ring over absolute output offsets: the last 32 KiB
  alignas(16) T win[WB];                 // the window, in coordinates relative to its 4-byte aligned base
  alignas(8) uint16_t val[WB];           // per window byte: V_FINAL or the window offset of an equal byte
  uint16_t queue[2][WB];                 // pending bytes of this / the next level
  alignas(8) uint2 tent[WT + 1];         // per token: x = token, y = byte offset in the window | (distance - 1) << 16;
                                         // [0] = tail of the token carried over from the previous window
  uint32_t bitmap[WB / 32];              // token-start bits over the window's bytes
  uint16_t wrank[WB / 32];               // token starts in the bitmap words before this one
  uint32_t qcnt[3];
  uint32_t hdr[SLAB_HDR_WORDS];
  uint32_t segstart[NL + 1];             // flat index of the first token of every list of the current slab
  uint32_t segptr[NL];                   // word offset of that token in the slab
  uint32_t crc_tab[256];
  uint32_t x16[WB / 16 + 4];             // x^(8 * 16 k) mod P: shifts a CRC over k 16-byte units
  uint32_t crcw[NWARP];
  uint32_t wscan[NWARP], wscan2[NWARP];
  unsigned long long wsum[NWARP][2];
  uint32_t member;
  int fail;
  uint32_t carry_len, carry_tok;
  uint32_t crc;
};
typedef SmemT<uint8_t> Smem;

__device__ __forceinline__ uint32_t tok_len(uint32_t t) { return (t & TOK_MATCH) ? (t & 255u) + 3u : 1u + ((t >> 30) & 1u); }

// CRC-32 of hist[a, a+m): every thread takes one contiguous slice; slices are merged pairwise with
// x^(8 len) shifts (the per-level shift is the square of the previous one).  All threads must call.
__device__ inline void crc_window(Smem &sm, uint32_t a, uint32_t m, int tid) {
  const uint32_t seg = (m + NT - 1) / NT;
  uint32_t lo = seg * tid, hi = lo + seg;
  if (lo > m) lo = m;
  if (hi > m) hi = m;
  uint32_t c = 0xffffffffu;
  for (uint32_t p = lo; p < hi; p++) c = (c >> 8) ^ sm.crc_tab[(c ^ sm.hist[(a + p) & HMASK]) & 0xff];
  c ^= 0xffffffffu;
  if (lo == hi) c = 0;
  uint32_t len = hi - lo;
  uint32_t shift = crc_x8n(seg);
  uint32_t *s_c = reinterpret_cast<uint32_t *>(sm.tent), *s_l = s_c + NT;   // the token array is dead by now
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
  unsigned long long acc_a, acc_w;        // per thread: Adler sum d, sum i*d over the bytes it flushed
  uint32_t carry_len, carry_tok;          // tail of the token that straddled the previous window end (as a token)
};

// One window: the slab's tokens [f, f + n) in flat order (n <= WT); consumes as many as fit,
// returns the number consumed (0xffffffff = the member must go to the sequential kernel).  A
// pending carry is flushed first.  All threads must call; the result is uniform.
template <typename T>
__device__ inline uint32_t resolve_window(T *__restrict__ out, int fmt, const uint32_t *__restrict__ slab, uint32_t f, uint32_t n,
                                          RState &rs, SmemT<T> &sm, int tid) {
  typedef typename Vec4<T>::type V4;
  constexpr uint32_t UNIT = 16 / sizeof(T);           // elements per 16-byte flush unit
  const int lane = tid & 31, warp = tid >> 5;
  const uint32_t pos = rs.pos;
  const uint32_t mis = pos & 3u, P4 = pos - mis;      // the window's coordinates start at the aligned base
  const uint32_t carry_len = rs.carry_len, carry_tok = rs.carry_tok;
  // ---- 1. tokens and their offsets
  const uint32_t tpt = (n + NT - 1) / NT;             // consecutive tokens per thread (<= TPT)
  uint32_t tk[TPT], ln[TPT];
  uint32_t mine = 0;
  {
    // the list that holds the thread's first token: last j with segstart[j] <= g
    const uint32_t g0 = f + tid * tpt;
    uint32_t j = 0;
    if (tid * tpt < n) {
#pragma unroll
      for (int stp = NL / 2; stp; stp >>= 1)
        if (sm.segstart[j + stp] <= g0) j += stp;
    }
    const uint32_t last = tid * tpt + tpt;             // one past the thread's last token
    if (last <= n && f + last <= sm.segstart[j + 1]) {  // common: all of them in one list
      const uint32_t *src = slab + sm.segptr[j] + (g0 - sm.segstart[j]);
#pragma unroll
      for (int q = 0; q < TPT; q++) {
        tk[q] = (uint32_t)q < tpt ? __ldg(src + q) : 0u;
        ln[q] = (uint32_t)q < tpt ? tok_len(tk[q]) : 0u;
        mine += ln[q];
      }
    } else {
#pragma unroll
      for (int q = 0; q < TPT; q++) {
        const uint32_t idx = tid * tpt + q;
        const bool have = (uint32_t)q < tpt && idx < n;
        tk[q] = 0u;
        if (have) {
          const uint32_t g = f + idx;
          while (g >= sm.segstart[j + 1]) j++;
          tk[q] = __ldg(slab + sm.segptr[j] + (g - sm.segstart[j]));
        }
        ln[q] = have ? tok_len(tk[q]) : 0u;
        mine += ln[q];
      }
    }
  }
  uint32_t x = mine;
#pragma unroll
  for (int sft = 1; sft < 32; sft <<= 1) {
    const uint32_t u = __shfl_up_sync(TBZ_FULL, x, sft);
    if (lane >= sft) x += u;
  }
  if (lane == 31) sm.wscan[warp] = x;
  if (tid < (int)(WB / 32)) sm.bitmap[tid] = 0;
  if (tid == 0) { sm.carry_len = 0; sm.qcnt[0] = 0; sm.qcnt[1] = 0; sm.qcnt[2] = 0; }
  __syncthreads();
  uint32_t off = mis + carry_len, total = mis + carry_len;
#pragma unroll
  for (int w = 0; w < NWARP; w++) { const uint32_t c = sm.wscan[w]; if (w < warp) off += c; total += c; }
  const uint32_t wend = total < WB ? total : WB;      // window = [mis, wend) in aligned coordinates
  uint32_t st = off + x - mine;
  uint32_t used = 0;
  bool bad = false;
#pragma unroll
  for (int q = 0; q < TPT; q++) {
    const uint32_t idx = tid * tpt + q;
    const bool have = (uint32_t)q < tpt && idx < n;
    if (have && st < WB) {
      used++;
      sm.tent[1 + idx] = make_uint2(tk[q], st | (((tk[q] >> 8) & 0x7fffu) << 16));
      atomicOr(&sm.bitmap[st >> 5], 1u << (st & 31u));
      if (tk[q] & TOK_MATCH) {
        const uint32_t d = ((tk[q] >> 8) & 0x7fffu) + 1u;
        if (d > P4 + st) bad = true;                              // deflate.lisp:343-345
        if (st + ln[q] > WB) { sm.carry_len = st + ln[q] - WB; sm.carry_tok = tk[q] & 0xffffff00u; }   // the straddler
      } else if (st + ln[q] > WB) { sm.carry_len = 1; sm.carry_tok = (tk[q] >> 8) & 255u; }          // second of two literals

    }
    st += ln[q];
  }
  if (bad) sm.fail = 1;
  if (tid == 0) {
    sm.tent[0] = make_uint2(carry_tok, mis | (((carry_tok >> 8) & 0x7fffu) << 16));
    if (carry_len) atomicOr(&sm.bitmap[0], 1u << mis);
  }
  uint32_t nused = n;
  if (total > WB) {                                   // rare: count the tokens that start inside the window
#pragma unroll
    for (int sft = 16; sft; sft >>= 1) used += __shfl_xor_sync(TBZ_FULL, used, sft);
    if (lane == 0) sm.wscan2[warp] = used;
  }
  __syncthreads();
  if (total > WB) {
    nused = 0;
#pragma unroll
    for (int w = 0; w < NWARP; w++) nused += sm.wscan2[w];
  }
  // rank directory over the bitmap words
  if (warp == 0) {
    uint32_t c[WB / 1024], s = 0;
#pragma unroll
    for (int i = 0; i < (int)(WB / 1024); i++) { c[i] = __popc(sm.bitmap[lane * (WB / 1024) + i]); s += c[i]; }
    uint32_t y = s;
#pragma unroll
    for (int sft = 1; sft < 32; sft <<= 1) {
      const uint32_t u = __shfl_up_sync(TBZ_FULL, y, sft);
      if (lane >= sft) y += u;
    }
    uint32_t run = y - s;
#pragma unroll
    for (int i = 0; i < (int)(WB / 1024); i++) { sm.wrank[lane * (WB / 1024) + i] = (uint16_t)run; run += c[i]; }
  }
  __syncthreads();
  if (sm.fail) return 0xffffffffu;
  const uint32_t adj = carry_len ? 1u : 0u;           // rank 1 is the carry pseudo-token (index 0) if there is one
  // ---- 2. every byte: literal, final history, or pending
  for (uint32_t rb = 0; rb < wend; rb += 4u * NT) {   // uniform trip count: the warp votes inside
    const uint32_t r0 = rb + 4u * tid;
    const uint32_t bw = sm.bitmap[(r0 >> 5) & (WB / 32 - 1)];
    uint32_t ti = sm.wrank[(r0 >> 5) & (WB / 32 - 1)] + __popc(bw & (0xffffffffu >> (31u - (r0 & 31u)))) - adj;
    const uint32_t nib = bw >> (r0 & 31u);
    uint32_t v01 = 0, v23 = 0, npend = 0, pmask = 0;
    T el[4];
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const uint32_t r = r0 + j;
      if (j) ti += (nib >> j) & 1u;
      // the same instructions for every byte: the history read is issued unconditionally (the ring
      // index is always valid) and selected afterwards; only the period of an overlapping match branches
      const uint2 e = sm.tent[ti < WT ? ti : WT];         // (a byte outside the window may have no token: any entry will do)
      const bool valid = r >= mis && r < wend;
      const bool ism = (e.x & TOK_MATCH) != 0;
      const uint32_t o = r - (e.y & 0xffffu), d = (e.y >> 16) + 1u;
      uint32_t back = d;
      if (__builtin_expect(valid && ism && o >= d, 0)) back = o - o % d + d;   // overlapping match: read through the period
      const T hv = sm.hist[(P4 + r - back) & HMASK];
      const bool fin = back + mis > r;                    // the source is below the window: final
      const T lit = (T)((e.x >> (8u * (o & 1u))) & 255u); // one or two literals in a token
      const T byte = ism ? hv : lit;
      const bool pend = valid && ism && !fin;
      const uint32_t v = pend ? r - back : V_FINAL;
      npend += pend ? 1u : 0u;
      pmask |= (pend ? 1u : 0u) << j;
      el[j] = byte;
      if (j < 2) v01 |= v << (16 * j); else v23 |= v << (16 * (j - 2));
    }
    if (r0 < wend) {
      if (sizeof(T) == 1) *reinterpret_cast<uint32_t *>(&sm.win[r0]) = (uint32_t)el[0] | ((uint32_t)el[1] << 8) | ((uint32_t)el[2] << 16) | ((uint32_t)el[3] << 24);
      else *reinterpret_cast<uint2 *>(&sm.win[r0]) = make_uint2((uint32_t)el[0] | ((uint32_t)el[1] << 16), (uint32_t)el[2] | ((uint32_t)el[3] << 16));
      *reinterpret_cast<uint2 *>(&sm.val[r0]) = make_uint2(v01, v23);
    }
    // queue the pending bytes: one shared-memory atomic per warp
    uint32_t incl = npend;
#pragma unroll
    for (int sft = 1; sft < 32; sft <<= 1) {
      const uint32_t u = __shfl_up_sync(TBZ_FULL, incl, sft);
      if (lane >= sft) incl += u;
    }
    uint32_t base = 0;
    if (lane == 31 && incl) base = atomicAdd(&sm.qcnt[0], incl);
    base = __shfl_sync(TBZ_FULL, base, 31);
    uint32_t qi = base + incl - npend;
#pragma unroll
    for (int j = 0; j < 4; j++)
      if (pmask & (1u << j)) sm.queue[0][qi++] = (uint16_t)(r0 + j);
  }
  // ---- 3. pointer jumping over the pending bytes
  for (uint32_t lvl = 0;; lvl++) {
    __syncthreads();
    const uint32_t qn = sm.qcnt[lvl % 3];
    if (!qn) break;
    if (tid == 0) sm.qcnt[(lvl + 2) % 3] = 0;
    const uint16_t *qin = sm.queue[lvl & 1];
    uint16_t *qout = sm.queue[(lvl + 1) & 1];
    for (uint32_t i = tid; i < qn; i += NT) {
      const uint32_t r = qin[i];
      const uint32_t s = sm.val[r];
      const uint32_t vs = *reinterpret_cast<volatile uint16_t *>(&sm.val[s]);
      if (vs == V_FINAL) {
        sm.win[r] = *reinterpret_cast<volatile T *>(&sm.win[s]);
        __threadfence_block();
        *reinterpret_cast<volatile uint16_t *>(&sm.val[r]) = (uint16_t)V_FINAL;
      } else {
        sm.val[r] = (uint16_t)vs;                      // equal bytes: adopt the source's pointer
        qout[atomicAdd(&sm.qcnt[(lvl + 1) % 3], 1u)] = (uint16_t)r;
      }
    }
  }
  // ---- 4. append the window to the history ring
  for (uint32_t r0 = 4u * tid; r0 < wend; r0 += 4u * NT) {
    if (r0 >= mis && r0 + 4u <= wend) *reinterpret_cast<V4 *>(&sm.hist[(P4 + r0) & HMASK]) = *reinterpret_cast<const V4 *>(&sm.win[r0]);
    else {
#pragma unroll
      for (int j = 0; j < 4; j++)
        if (r0 + j >= mis && r0 + j < wend) sm.hist[(P4 + r0 + j) & HMASK] = sm.win[r0 + j];
    }
  }
  __syncthreads();
  const uint32_t wsize = wend - mis;
  // ---- 5. flush complete 16-byte units, fold them into the checksum
  const bool aligned_out = (((uintptr_t)out) & 15) == 0;
  if (sizeof(T) == 1 && fmt == TBZ_GZIP && !aligned_out) crc_window(reinterpret_cast<Smem &>(sm), pos, wsize, tid);
  if (aligned_out) {
    if (rs.flushed & (UNIT - 1u)) {                    // a chunk of a split member starts inside a unit: element-wise head
      uint32_t upto = (rs.flushed + UNIT - 1u) & ~(UNIT - 1u);
      if (upto > pos + wsize) upto = pos + wsize;
      if (rs.flushed + tid < upto) out[rs.flushed + tid] = sm.hist[(rs.flushed + tid) & HMASK];
      rs.flushed = upto;
    }
    const uint32_t upto = (pos + wsize) & ~(UNIT - 1u);
    uint32_t myc = 0;                                  // gzip: CRCs of this thread's units, shifted to the end of the flushed range
    for (uint32_t p = rs.flushed + UNIT * tid; p < upto; p += UNIT * NT) {
      const uint4 v = *reinterpret_cast<const uint4 *>(&sm.hist[p & HMASK]);
      *reinterpret_cast<uint4 *>(out + p) = v;
      if (sizeof(T) == 1 && fmt == TBZ_GZIP) {
        // crc(A || B) = crc(A) * x^(8 |B|) + crc(B) for finalized CRCs: one table CRC per unit, one
        // multiplication by the power for the bytes that follow it, XOR over all units
        const uint32_t w4[4] = {v.x, v.y, v.z, v.w};
        uint32_t c = 0xffffffffu;
#pragma unroll
        for (int q = 0; q < 4; q++)
#pragma unroll
          for (int b8 = 0; b8 < 4; b8++) c = (c >> 8) ^ sm.crc_tab[(c ^ (w4[q] >> (8 * b8))) & 0xff];
        myc ^= crc_mulmod(sm.x16[(upto - p - 16u) >> 4], c ^ 0xffffffffu);
      }
      if (sizeof(T) == 1 && fmt == TBZ_ZLIB) {
        uint32_t sd = __dp4a(v.x, 0x01010101u, 0u); sd = __dp4a(v.y, 0x01010101u, sd);
        sd = __dp4a(v.z, 0x01010101u, sd); sd = __dp4a(v.w, 0x01010101u, sd);
        uint32_t wj = __dp4a(v.x, 0x03020100u, 0u); wj = __dp4a(v.y, 0x07060504u, wj);
        wj = __dp4a(v.z, 0x0b0a0908u, wj); wj = __dp4a(v.w, 0x0f0e0d0cu, wj);
        rs.acc_a += sd;
        rs.acc_w += (unsigned long long)p * sd + wj;
      }
    }
    if (sizeof(T) == 1 && fmt == TBZ_GZIP) {
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
      const uint32_t d = sm.hist[p & HMASK];
      out[p] = (T)d;
      rs.acc_a += d; rs.acc_w += (unsigned long long)p * d;
    }
    rs.flushed = pos + wsize;
  }
  if (__builtin_expect((rs.acc_w >> 62) != 0, 0)) rs.acc_w %= TBZ_ADLER_MOD;
  rs.pos = pos + wsize;
  rs.carry_len = sm.carry_len; rs.carry_tok = sm.carry_tok;
  __syncthreads();
  return nused;
}

// Every window of one token stream (a member, or one chunk of a split member).  rs.pos / rs.flushed
// hold the absolute output offset the stream starts at.  Returns false when the caller must fall back.
template <typename T>
__device__ inline bool resolve_stream(T *__restrict__ out, int fmt, const P1Rec &rec, const uint32_t *__restrict__ slabs,
                                      RState &rs, SmemT<T> &sm, int tid) {
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
    while (f < total) {
      const uint32_t n = total - f < WT ? total - f : WT;
      const uint32_t used = resolve_window<T>(out, fmt, slab, f, n, rs, sm, tid);
      if (used == 0xffffffffu || used == 0) return false;
      f += used;
    }
    __syncthreads();
  }
  while (rs.carry_len) {                   // tail of a token that straddled the last window
    if (resolve_window<T>(out, fmt, nullptr, 0, 0, rs, sm, tid) == 0xffffffffu) return false;
  }
  if (sm.fail) return false;
  if (rs.flushed + tid < rs.pos) {         // the last partial 16-byte unit
    const uint32_t p = rs.flushed + tid;
    const uint32_t d = sm.hist[p & HMASK];
    out[p] = (T)d;
    rs.acc_a += d; rs.acc_w += (unsigned long long)p * d;
  }
  return true;
}

__device__ inline bool resolve_member(const DMember &mem, int fmt, const P1Rec &rec, const uint32_t *__restrict__ slabs,
                                      tbz_result &res, Smem &sm, int tid) {
  const int lane = tid & 31, warp = tid >> 5;
  RState rs;
  rs.pos = 0; rs.flushed = 0; rs.acc_a = 0; rs.acc_w = 0; rs.carry_len = 0; rs.carry_tok = 0;
  if (!resolve_stream<uint8_t>(mem.out, fmt, rec, slabs, rs, sm, tid)) return false;
  const uint32_t pos = rs.pos;
  if (pos != rec.out_len) return false;
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
  } else if (fmt == TBZ_GZIP) {
    __syncthreads();
    uint32_t c = sm.crc;
    if (rs.flushed < pos) {                  // the last partial unit (uniform: every thread computes the same value)
      uint32_t t = 0xffffffffu;
      for (uint32_t p = rs.flushed; p < pos; p++) t = (t >> 8) ^ sm.crc_tab[(t ^ sm.hist[p & HMASK]) & 0xff];
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
    if (t != ck) return false;
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

}  // namespace tbzres
