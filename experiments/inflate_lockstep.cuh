// inflate_lockstep.cuh — phase two of the batched fast path for byte members: LZ77 resolution of a
// token stream (deflate.lisp:244-359 `copy-history`) by lanes that walk the tokens in lock step.
//
// One CTA per member, the 32 KiB deflate history as a ring in shared memory.  A window is the next
// <= WB output bytes; lane i of the CTA owns bytes [C i, C i + C) of it (C = 16) and produces them
// one per iteration, exactly as a sequential decoder would: it walks its tokens, a literal is a byte,
// a match byte is read from `distance` bytes back.  All lanes of a warp are at the same iteration t,
// and the warps of the CTA meet at a barrier every KPH iterations, so at iteration t a source byte
//   * below the window                      is final history                    -> read from the ring
//   * at offset u < t of a chunk of this warp (u < the last barrier's t for another warp's chunk)
//                                           has been produced                   -> read from the window
//   * anywhere else in the window           does not exist yet                  -> the byte becomes a
//                                           POINTER to its source (an equal byte), resolved below
// The window holds 16-bit symbols: FINAL | byte, or the window offset of an equal byte; a produced
// source that is itself a pointer is simply copied (the pointer is adopted).  Per byte this costs a
// few instructions and one or two shared-memory accesses, there is no per-byte token search, and a
// lane reads its tokens sequentially.  Pointers (about a quarter of the bytes on text) are then
// resolved by pointer jumping over a dense queue (chains halve per level), the window is packed to
// bytes, appended to the ring and flushed with 16-byte stores; Adler-32 / CRC-32 are folded in.
// The trailer is checked as zlib.lisp:80-96 / gzip.lisp:82-106 do; any disagreement sends the member
// to the sequential kernel, which owns the verdict rules.
#pragma once
#include "tbz_device.cuh"
#include "inflate_decode.cuh"

namespace tbzls {

using tbzfast::NL;
using tbzfast::NO_SLAB;
using tbzfast::P1Rec;
using tbzfast::SLAB_HDR_WORDS;
using tbzfast::SLAB_WORDS;
using tbzfast::SlabHdr;
using tbzfast::TOKCAP;
using tbzfast::TOK_MATCH;

constexpr int NT = 256;
constexpr int NWARP = NT / 32;
constexpr uint32_t HIST = 32768u, HMASK = HIST - 1u;
constexpr bool CRC_SEPARATE = false;    // CRC-32 inside the kernel
#ifndef TBZ_LS_C
#define TBZ_LS_C 16
#endif
#ifndef TBZ_LS_K
#define TBZ_LS_K 4
#endif
#ifndef TBZ_LS_TPT
#define TBZ_LS_TPT 4
#endif
constexpr int C = TBZ_LS_C;                 // bytes per lane and window
constexpr int LOGC = C == 8 ? 3 : C == 16 ? 4 : 5;
static_assert(C == 8 || C == 16 || C == 32, "chunk size");
constexpr int KPH = TBZ_LS_K;               // iterations between two CTA barriers
constexpr int TPT = TBZ_LS_TPT;             // tokens per thread and window
constexpr uint32_t WB = (uint32_t)NT * C;   // window bytes (including the <= 15 bytes of alignment lead-in)
constexpr uint32_t WT = (uint32_t)NT * TPT; // window tokens
constexpr uint32_t CSTR = C + 2;            // symbols per chunk in shared memory: an odd number of 32-bit words, so
                                            // the lanes of a warp store to 32 different banks
constexpr uint32_t FINAL = 0x8000u;
static_assert(((CSTR * 2 / 4) & 1) == 1, "chunk stride must be an odd number of words");
static_assert(WB <= 0x8000u, "window offsets must fit under the FINAL flag");

struct Smem {
  alignas(16) uint8_t hist[HIST];          // ring over absolute output offsets: the last 32 KiB
  alignas(16) uint16_t win[NT * CSTR];     // the window's symbols, chunk i at [CSTR i, CSTR i + C)
  alignas(8) uint2 tent[WT + 2];           // per token: x = token, y = byte offset in the window | (distance - 1) << 16;
                                           // [0] = tail of the token carried over from the previous window
  uint16_t queue[2][WB];                   // pointer bytes of this / the next level
  uint16_t ctok[NT + 1];                   // per chunk: index in tent of the token that covers its first byte
  uint32_t qcnt[3];
  uint32_t hdr[SLAB_HDR_WORDS];
  uint32_t segstart[NL + 1];               // flat index of the first token of every list of the current slab
  uint32_t segptr[NL];                     // word offset of that token in the slab
  uint32_t crc_tab[256];
  uint32_t x16[WB / 16 + 4];               // x^(8 * 16 k) mod P: shifts a CRC over k 16-byte units
  uint32_t crcw[NWARP];
  uint32_t wscan[NWARP], wscan2[NWARP];
  unsigned long long wsum[NWARP][2];
  uint32_t member;
  int fail;
  uint32_t carry_len, carry_tok;
  uint32_t crc;
};

__device__ __forceinline__ uint32_t tok_len(uint32_t t) { return (t & TOK_MATCH) ? (t & 255u) + 3u : 1u + ((t >> 30) & 1u); }
__device__ __forceinline__ uint32_t widx(uint32_t r) { return r + 2u * (r >> LOGC); }   // window offset -> index in win[]

// CRC-32 of hist[a, a+m): every thread takes one contiguous slice; slices are merged pairwise with
// x^(8 len) shifts.  Only used when the output pointer is not 16-byte aligned.  All threads must call.
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
// pending carry is produced first.  All threads must call; the result is uniform.
__device__ inline uint32_t resolve_window(uint8_t *__restrict__ out, int fmt, const uint32_t *__restrict__ slab, uint32_t f, uint32_t n,
                                          RState &rs, Smem &sm, int tid) {
  const int lane = tid & 31, warp = tid >> 5;
  const uint32_t pos = rs.pos;
  const uint32_t mis = pos & 15u, P16 = pos - mis;    // the window's coordinates start at the 16-byte aligned base
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
      // the chunks whose first byte this token covers
      const uint32_t e = st + ln[q] < WB ? st + ln[q] : WB;
      for (uint32_t b = (st + C - 1u) & ~(uint32_t)(C - 1); b < e; b += C) sm.ctok[b >> LOGC] = (uint16_t)(1u + idx);
      if (tk[q] & TOK_MATCH) {
        const uint32_t d = ((tk[q] >> 8) & 0x7fffu) + 1u;
        if (d > P16 + st) bad = true;                             // deflate.lisp:343-345
        if (st + ln[q] > WB) { sm.carry_len = st + ln[q] - WB; sm.carry_tok = tk[q] & 0xffffff00u; }   // the straddler
      } else if (st + ln[q] > WB) { sm.carry_len = 1; sm.carry_tok = (tk[q] >> 8) & 255u; }          // second of two literals
    }
    st += ln[q];
  }
  if (bad) sm.fail = 1;
  if (tid == 0) {
    sm.tent[0] = make_uint2(carry_tok, mis | (((carry_tok >> 8) & 0x7fffu) << 16));
    // the chunk that holds the first byte of the window, and those the carried tail reaches into
    sm.ctok[mis >> LOGC] = carry_len ? 0u : 1u;
    for (uint32_t b = (mis + C) & ~(uint32_t)(C - 1); b < mis + carry_len; b += C) sm.ctok[b >> LOGC] = 0u;
  }
  uint32_t nused = n;
  if (total > WB) {                                   // count the tokens that start inside the window
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
  if (sm.fail) return 0xffffffffu;
  // ---- 2. lock-step production of the window's symbols
  const uint32_t r0 = (uint32_t)tid * C;
  uint16_t *my = sm.win + (uint32_t)tid * CSTR;
  uint32_t pmask = 0;
  {
    const uint32_t rs0 = r0 < mis ? mis : r0;          // the first byte this lane produces
    uint32_t ti = 0, rem = 0, lits = 0;
    int srel = 0;                                      // window offset of the next source byte (below mis: history)
    bool ism = false;
    if (rs0 < wend && rs0 < r0 + C) {
      ti = sm.ctok[tid];
      const uint2 e = sm.tent[ti];
      const uint32_t o = rs0 - (e.y & 0xffffu);        // offset inside the token
      const uint32_t len = ti ? tok_len(e.x) : carry_len;
      const uint32_t d = (e.y >> 16) + 1u;
      ism = (e.x & TOK_MATCH) != 0;
      rem = len - o;
      lits = e.x >> (8u * (o & 1u));
      uint32_t back = d;
      if (ism && o >= d) back = o - o % d + d;         // overlapping match: any whole number of periods back is an equal byte
      srel = (int)rs0 - (int)back;
    }
#pragma unroll
    for (int t = 0; t < C; t++) {
      if (KPH < C && t && (t % KPH) == 0) __syncthreads();
      const int t0 = KPH < C ? t - t % KPH : 0;        // what the other warps are known to have produced
      const uint32_t r = r0 + (uint32_t)t;
      const bool act = r >= mis && r < wend;
      if (act && rem == 0) {                           // the next token starts here
        ti++;
        const uint2 e = sm.tent[ti];
        ism = (e.x & TOK_MATCH) != 0;
        rem = tok_len(e.x);
        lits = e.x;
        srel = (int)r - (int)((e.y >> 16) + 1u);
      }
      uint32_t sym = FINAL | (lits & 255u);
      if (act && ism) {
        if (srel < (int)mis) sym = FINAL | sm.hist[(P16 + (uint32_t)srel) & HMASK];
        else {
          const uint32_t s = (uint32_t)srel;
          const uint32_t lim = ((s >> (LOGC + 5)) == (uint32_t)warp) ? (uint32_t)t : (uint32_t)t0;
          sym = s;
          if ((s & (uint32_t)(C - 1)) < lim) sym = sm.win[widx(s)];
        }
      }
      if (act) {
        my[t] = (uint16_t)sym;
        if (!(sym & FINAL)) pmask |= 1u << t;
        rem--;
        srel++;
        lits >>= 8;
      }
      __syncwarp();
    }
  }
  // ---- 3. queue the pointer bytes (one shared-memory atomic per warp), resolve them by pointer jumping
  {
    const uint32_t npend = __popc(pmask);
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
    while (pmask) {
      const uint32_t j = __ffs(pmask) - 1u;
      pmask &= pmask - 1u;
      sm.queue[0][qi++] = (uint16_t)(r0 + j);
    }
  }
  for (uint32_t lvl = 0;; lvl++) {
    __syncthreads();
    const uint32_t qn = sm.qcnt[lvl % 3];
    if (!qn) break;
    if (tid == 0) sm.qcnt[(lvl + 2) % 3] = 0;
    const uint16_t *qin = sm.queue[lvl & 1];
    uint16_t *qout = sm.queue[(lvl + 1) & 1];
    for (uint32_t i = tid; i < qn; i += NT) {
      const uint32_t r = qin[i];
      volatile uint16_t *mine_p = reinterpret_cast<volatile uint16_t *>(&sm.win[widx(r)]);
      const uint32_t s = *mine_p;
      const uint32_t vs = *reinterpret_cast<volatile uint16_t *>(&sm.win[widx(s)]);
      *mine_p = (uint16_t)vs;                            // the byte itself, or (equal bytes) the source's pointer
      if (!(vs & FINAL)) qout[atomicAdd(&sm.qcnt[(lvl + 1) % 3], 1u)] = (uint16_t)r;
    }
  }
  // ---- 4. pack the window to bytes and append it to the history ring
  if (r0 >= mis && r0 + C <= wend) {
    const uint32_t *w = reinterpret_cast<const uint32_t *>(my);
    uint32_t o4[C / 4];
#pragma unroll
    for (int k = 0; k < C / 4; k++) o4[k] = __byte_perm(w[2 * k], w[2 * k + 1], 0x6420);
    uint8_t *dst = &sm.hist[(P16 + r0) & HMASK];
    if (C == 8) *reinterpret_cast<uint2 *>(dst) = make_uint2(o4[0], o4[1]);
    else {
#pragma unroll
      for (int k = 0; k < C / 16; k++) *reinterpret_cast<uint4 *>(dst + 16 * k) = make_uint4(o4[4 * k], o4[4 * k + 1], o4[4 * k + 2], o4[4 * k + 3]);
    }
  } else {
    for (uint32_t j = 0; j < (uint32_t)C; j++)
      if (r0 + j >= mis && r0 + j < wend) sm.hist[(P16 + r0 + j) & HMASK] = (uint8_t)my[j];
  }
  __syncthreads();
  const uint32_t wsize = wend - mis;
  // ---- 5. flush complete 16-byte units, fold them into the checksum
  const bool aligned_out = (((uintptr_t)out) & 15) == 0;
  if (fmt == TBZ_GZIP && !aligned_out) crc_window(sm, pos, wsize, tid);
  if (aligned_out) {
    const uint32_t upto = (pos + wsize) & ~15u;
    uint32_t myc = 0;                                  // gzip: CRCs of this thread's units, shifted to the end of the flushed range
    for (uint32_t p = rs.flushed + 16u * tid; p < upto; p += 16u * NT) {
      const uint4 v = *reinterpret_cast<const uint4 *>(&sm.hist[p & HMASK]);
      *reinterpret_cast<uint4 *>(out + p) = v;
      if (fmt == TBZ_GZIP) {
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
      if (fmt == TBZ_ZLIB) {
        uint32_t sd = __dp4a(v.x, 0x01010101u, 0u); sd = __dp4a(v.y, 0x01010101u, sd);
        sd = __dp4a(v.z, 0x01010101u, sd); sd = __dp4a(v.w, 0x01010101u, sd);
        uint32_t wj = __dp4a(v.x, 0x03020100u, 0u); wj = __dp4a(v.y, 0x07060504u, wj);
        wj = __dp4a(v.z, 0x0b0a0908u, wj); wj = __dp4a(v.w, 0x0f0e0d0cu, wj);
        rs.acc_a += sd;
        rs.acc_w += (unsigned long long)p * sd + wj;
      }
    }
    if (fmt == TBZ_GZIP) {
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
      out[p] = (uint8_t)d;
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
    while (f < total) {
      const uint32_t n = total - f < WT ? total - f : WT;
      const uint32_t used = resolve_window(out, fmt, slab, f, n, rs, sm, tid);
      if (used == 0xffffffffu || used == 0) return false;
      f += used;
    }
    __syncthreads();
  }
  while (rs.carry_len) {                   // tail of a token that straddled the last window
    if (resolve_window(out, fmt, nullptr, 0, 0, rs, sm, tid) == 0xffffffffu) return false;
  }
  if (sm.fail) return false;
  if (rs.flushed + tid < rs.pos) {         // the last partial 16-byte unit
    const uint32_t p = rs.flushed + tid;
    const uint32_t d = sm.hist[p & HMASK];
    out[p] = (uint8_t)d;
    rs.acc_a += d; rs.acc_w += (unsigned long long)p * d;
  }
  return true;
}

__device__ inline bool resolve_member(const DMember &mem, int fmt, const P1Rec &rec, const uint32_t *__restrict__ slabs,
                                      tbz_result &res, Smem &sm, int tid) {
  const int lane = tid & 31, warp = tid >> 5;
  RState rs;
  rs.pos = 0; rs.flushed = 0; rs.acc_a = 0; rs.acc_w = 0; rs.carry_len = 0; rs.carry_tok = 0;
  if (!resolve_stream(mem.out, fmt, rec, slabs, rs, sm, tid)) return false;
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

}  // namespace tbzls
