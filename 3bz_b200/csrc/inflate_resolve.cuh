// inflate_resolve.cuh — phase two of the batched fast path: LZ77 resolution of a token stream
// (deflate.lisp:244-359 `copy-history`, restated for a whole CTA).
//
// One CTA per member.  The member's output is produced in windows of up to WB bytes that live in
// a 64 KiB ring in shared memory (the last 32 KiB of it are the deflate history):
//   1. the next <= WT tokens are fetched in flat order (two per thread) from the member's slabs
//   2. a CTA prefix sum over their lengths gives every token its offset in the window; the token
//      that straddles the window end is split and its tail carried into the next window
//   3. token indices are scattered to their start offsets and a prefix-max turns that into a
//      byte -> token map
//   4. every output byte is resolved *by address arithmetic only*: follow byte -> token ->
//      (byte - distance) while the source still lies inside the window (overlapping matches go
//      through their period); the chase ends at a literal token or at a byte below the window,
//      which is final and sits in the ring.  No byte written in this window is read in this window,
//      so all 256 threads work on 8 bytes each without any ordering between them
//   5. the window is flushed to global memory with 16-byte stores; Adler-32 is folded in with
//      dp4a as s1 = 1 + sum d, s2 = N + N sum d - sum i d_i (order independent per thread)
// CRC-32 (gzip) is a thread-parallel pass per window with x^(8 len) combines.  The trailer is then
// checked as zlib.lisp:80-96 / gzip.lisp:82-106 do; any disagreement sends the member to the
// sequential kernel, which owns the verdict rules.
#pragma once
#include "tbz_device.cuh"
#include "inflate_decode.cuh"

namespace tbzres {

using tbzfast::NO_SLAB;
using tbzfast::NT;
using tbzfast::NWARP;
using tbzfast::P1Rec;
using tbzfast::SLAB_WORDS;
using tbzfast::SlabHdr;
using tbzfast::TOKCAP;
using tbzfast::TOK_EOB;
using tbzfast::TOK_MATCH;

constexpr uint32_t RING = 65536u, RMASK = RING - 1u;
constexpr uint32_t WB = 2048;            // window bytes (8 per thread)
constexpr uint32_t WT = 2 * NT;          // window tokens (2 per thread)

struct Smem {
  alignas(16) uint8_t ring[RING];
  uint32_t toks[WT + 1];                 // [0] = tail of the match carried over from the previous window
  uint16_t tstart[WT + 2];
  alignas(16) uint16_t bytemap[WB];      // byte -> token index + 1; doubles as scratch of the crc tree
  uint16_t tb[NT];
  uint16_t g0[NT];
  uint32_t crc_tab[256];
  uint32_t wscan[NWARP], wscan2[NWARP];
  unsigned long long wsum[NWARP][2];
  uint32_t member;
  int fail;
  uint32_t carry_len, carry_dist;
  uint32_t crc;
};

__device__ __forceinline__ uint32_t tok_len(uint32_t t) {
  return (t & TOK_MATCH) ? (t & 255u) + 3u : ((t & TOK_EOB) ? 0u : 1u);
}

// CRC-32 of ring[a, a+m): every thread takes one contiguous slice; slices are merged pairwise with
// x^(8 len) shifts (the per-level shift is the square of the previous one).  All threads must call.
__device__ inline void crc_window(Smem &sm, uint32_t a, uint32_t m, int tid) {
  const uint32_t seg = (m + NT - 1) / NT;
  uint32_t lo = seg * tid, hi = lo + seg;
  if (lo > m) lo = m;
  if (hi > m) hi = m;
  uint32_t c = 0xffffffffu;
  for (uint32_t p = lo; p < hi; p++) c = (c >> 8) ^ sm.crc_tab[(c ^ sm.ring[(a + p) & RMASK]) & 0xff];
  c ^= 0xffffffffu;
  if (lo == hi) c = 0;
  uint32_t len = hi - lo;
  uint32_t shift = crc_x8n(seg);
  uint32_t *s_c = reinterpret_cast<uint32_t *>(sm.bytemap), *s_l = s_c + NT;   // the map is dead by now
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

__device__ inline bool resolve_member(const DMember &mem, int fmt, const P1Rec &rec, const uint32_t *__restrict__ slabs,
                                      tbz_result &res, Smem &sm, int tid) {
  const int lane = tid & 31, warp = tid >> 5;
  uint8_t *out = mem.out;
  const bool out_aligned = (((uintptr_t)out) & 15) == 0;
  uint32_t pos = 0;                       // output bytes produced so far (window base)
  uint32_t flushed = 0;                   // output bytes already stored to global memory
  unsigned long long acc_a = 0, acc_w = 0;   // Adler: sum d, sum i*d over this thread's flushed bytes
  if (tid == 0) { sm.fail = 0; sm.carry_len = 0; sm.carry_dist = 1; sm.crc = 0; }
  uint32_t carry_len = 0, carry_dist = 1;   // tail of a match that straddled the previous window end (uniform)
  __syncthreads();
  for (uint32_t s = rec.first_slab; s != NO_SLAB;) {
    const uint32_t *slab = slabs + (size_t)s * SLAB_WORDS;
    const SlabHdr *sh = reinterpret_cast<const SlabHdr *>(slab);
    const uint32_t *lists = slab + sizeof(SlabHdr) / 4;
    s = sh->next;
    const uint32_t ntok = sh->ntokens;
    sm.tb[tid] = (uint16_t)sh->tb[tid];
    sm.g0[tid] = (uint16_t)(sh->gn[tid] & 0xffffu);
    __syncthreads();
    uint32_t f = 0;                       // next flat token of this slab
    // the tail carried over from the previous slab is flushed with this slab's first window; a
    // slab without tokens still needs one pass if a tail is pending
    while (f < ntok || carry_len) {
      if (tid == 0) sm.carry_len = 0;      // rewritten below by the thread that owns a straddling match
      // ---- 1. fetch two tokens per thread, flat order.  One binary search per warp for the list
      // that holds the warp's first token, then every thread walks forward from there.
      uint32_t tk[2], ln[2];
      {
        const uint32_t fw = f + 64 * warp;          // first flat token of this warp
        uint32_t j0 = 0;                            // last lane with tb[j0] <= fw
        if (fw < ntok) {
#pragma unroll
          for (int stp = NT / 2; stp; stp >>= 1)
            if (sm.tb[j0 + stp] <= fw) j0 += stp;
        }
#pragma unroll
        for (int q = 0; q < 2; q++) {
          const uint32_t fi = f + 2 * tid + q;
          tk[q] = TOK_EOB;
          if (fi < ntok) {
            uint32_t j = j0;
            while (j + 1 < NT && sm.tb[j + 1] <= fi) j++;
            tk[q] = lists[j * TOKCAP + sm.g0[j] + (fi - sm.tb[j])];
          }
          ln[q] = tok_len(tk[q]);
        }
      }
      // ---- 2. offsets inside the window
      uint32_t x = ln[0] + ln[1];
      const uint32_t mine = x;
#pragma unroll
      for (int sft = 1; sft < 32; sft <<= 1) {
        const uint32_t u = __shfl_up_sync(TBZ_FULL, x, sft);
        if (lane >= sft) x += u;
      }
      if (lane == 31) sm.wscan[warp] = x;
      __syncthreads();
      uint32_t off = carry_len, total = carry_len;
#pragma unroll
      for (int w = 0; w < NWARP; w++) { const uint32_t c = sm.wscan[w]; if (w < warp) off += c; total += c; }
      uint32_t st[2];
      st[0] = off + x - mine;
      st[1] = st[0] + ln[0];
      const uint32_t wsize = total < WB ? total : WB;
      // tokens that start inside the window are consumed by it
      uint32_t used = 0;
      bool bad = false;
#pragma unroll
      for (int q = 0; q < 2; q++) {
        const uint32_t fi = f + 2 * tid + q;
        const bool inc = fi < ntok && st[q] < WB;
        used += inc;
        const uint32_t idx = 1 + 2 * tid + q;
        sm.toks[idx] = tk[q];
        sm.tstart[idx] = (uint16_t)(st[q] < WB ? st[q] : WB);
        if (inc && (tk[q] & TOK_MATCH)) {
          const uint32_t d = ((tk[q] >> 8) & 0x7fffu) + 1u;
          if (d > pos + st[q]) bad = true;                       // deflate.lisp:343-345
          if (st[q] + ln[q] > WB) { sm.carry_len = st[q] + ln[q] - WB; sm.carry_dist = d; }   // the straddler
        }
      }
      if (bad) sm.fail = 1;
      if (tid == 0) {
        sm.toks[0] = TOK_MATCH | ((carry_dist - 1) << 8) | ((carry_len >= 3 ? carry_len : 3) - 3);
        sm.tstart[0] = 0;
      }
      // ---- 3. byte -> token map: scatter the token starts, prefix-max
      *reinterpret_cast<uint4 *>(&sm.bytemap[8 * tid]) = make_uint4(0, 0, 0, 0);
#pragma unroll
      for (int sft = 16; sft; sft >>= 1) used += __shfl_xor_sync(TBZ_FULL, used, sft);
      if (lane == 0) sm.wscan2[warp] = used;
      __syncthreads();
      uint32_t nused = 0;
#pragma unroll
      for (int w = 0; w < NWARP; w++) nused += sm.wscan2[w];
#pragma unroll
      for (int q = 0; q < 2; q++)
        if (ln[q] && st[q] < WB && f + 2 * tid + q < ntok) sm.bytemap[st[q]] = (uint16_t)(2 + 2 * tid + q);   // token index + 1
      if (tid == 0 && carry_len) sm.bytemap[0] = 1;
      __syncthreads();
      if (sm.fail) return false;
      uint32_t mp[8];
      {
        const uint4 v = *reinterpret_cast<const uint4 *>(&sm.bytemap[8 * tid]);
        mp[0] = v.x & 0xffffu; mp[1] = v.x >> 16; mp[2] = v.y & 0xffffu; mp[3] = v.y >> 16;
        mp[4] = v.z & 0xffffu; mp[5] = v.z >> 16; mp[6] = v.w & 0xffffu; mp[7] = v.w >> 16;
      }
#pragma unroll
      for (int i = 1; i < 8; i++) mp[i] = mp[i] > mp[i - 1] ? mp[i] : mp[i - 1];
      uint32_t run = mp[7];
#pragma unroll
      for (int sft = 1; sft < 32; sft <<= 1) {
        const uint32_t u = __shfl_up_sync(TBZ_FULL, run, sft);
        if (lane >= sft && u > run) run = u;
      }
      if (lane == 31) sm.wscan[warp] = run;
      uint32_t before = __shfl_up_sync(TBZ_FULL, run, 1);
      if (lane == 0) before = 0;
      __syncthreads();
#pragma unroll
      for (int w = 0; w < NWARP; w++) { const uint32_t c = sm.wscan[w]; if (w < warp && c > before) before = c; }
#pragma unroll
      for (int i = 0; i < 8; i++) if (before > mp[i]) mp[i] = before;
      *reinterpret_cast<uint4 *>(&sm.bytemap[8 * tid]) =
          make_uint4(mp[0] | (mp[1] << 16), mp[2] | (mp[3] << 16), mp[4] | (mp[5] << 16), mp[6] | (mp[7] << 16));
      __syncthreads();
      // ---- 4. resolve the window in four 512-byte sub-passes, two bytes per thread each.  A byte
      // whose source lies below the sub-pass start reads it from the ring (earlier sub-passes have
      // stored there already), so only sources inside the same 512 bytes are chased further.
#pragma unroll 1
      for (uint32_t sub = 0; sub < WB; sub += 512) {
        if (sub < wsize) {
          uint32_t rr[2], byte[2];
          bool open[2];
#pragma unroll
          for (int i = 0; i < 2; i++) {
            rr[i] = sub + 2 * tid + i;
            byte[i] = 0;
            open[i] = rr[i] < wsize;
          }
#pragma unroll
          for (int hop = 0; hop < 2; hop++) {
#pragma unroll
            for (int i = 0; i < 2; i++) {
              if (open[i]) {
                const uint32_t m = sm.bytemap[rr[i]];
                const uint32_t t = sm.toks[m - 1];
                if (!(t & TOK_MATCH)) { byte[i] = t & 255u; open[i] = false; }
                else {
                  const uint32_t d = ((t >> 8) & 0x7fffu) + 1u, o = rr[i] - sm.tstart[m - 1];
                  uint32_t back = d;
                  if (o >= d) back = o - o % d + d;             // overlapping match: read through the period
                  if (back + sub > rr[i]) { byte[i] = sm.ring[(pos + rr[i] - back) & RMASK]; open[i] = false; }
                  else rr[i] -= back;
                }
              }
            }
          }
#pragma unroll
          for (int i = 0; i < 2; i++) {
            while (open[i]) {                                   // rare: three or more hops
              const uint32_t m = sm.bytemap[rr[i]];
              const uint32_t t = sm.toks[m - 1];
              if (!(t & TOK_MATCH)) { byte[i] = t & 255u; break; }
              const uint32_t d = ((t >> 8) & 0x7fffu) + 1u, o = rr[i] - sm.tstart[m - 1];
              uint32_t back = d;
              if (o >= d) back = o - o % d + d;
              if (back + sub > rr[i]) { byte[i] = sm.ring[(pos + rr[i] - back) & RMASK]; break; }
              rr[i] -= back;
            }
          }
          const uint32_t r0 = sub + 2 * tid;
          if (r0 < wsize) sm.ring[(pos + r0) & RMASK] = (uint8_t)byte[0];
          if (r0 + 1 < wsize) sm.ring[(pos + r0 + 1) & RMASK] = (uint8_t)byte[1];
        }
        __syncthreads();
      }
      // ---- 5. flush complete 16-byte units, fold them into the checksum
      if (fmt == TBZ_GZIP) crc_window(sm, pos, wsize, tid);
      if (out_aligned) {
        const uint32_t upto = (pos + wsize) & ~15u;
        const uint32_t p = flushed + 16 * tid;
        if (p < upto) {
          const uint4 v = *reinterpret_cast<const uint4 *>(&sm.ring[p & RMASK]);
          *reinterpret_cast<uint4 *>(out + p) = v;
          if (fmt == TBZ_ZLIB) {
            uint32_t sd = __dp4a(v.x, 0x01010101u, 0u); sd = __dp4a(v.y, 0x01010101u, sd);
            sd = __dp4a(v.z, 0x01010101u, sd); sd = __dp4a(v.w, 0x01010101u, sd);
            uint32_t wj = __dp4a(v.x, 0x03020100u, 0u); wj = __dp4a(v.y, 0x07060504u, wj);
            wj = __dp4a(v.z, 0x0b0a0908u, wj); wj = __dp4a(v.w, 0x0f0e0d0cu, wj);
            acc_a += sd;
            acc_w += (unsigned long long)p * sd + wj;
          }
        }
        flushed = upto;                                        // at most WB + 15 bytes = 129 units per window
      } else {
        for (uint32_t p = pos + tid; p < pos + wsize; p += NT) {
          const uint32_t d = sm.ring[p & RMASK];
          out[p] = (uint8_t)d;
          acc_a += d; acc_w += (unsigned long long)p * d;
        }
        flushed = pos + wsize;
      }
      if (acc_w >> 62) acc_w %= TBZ_ADLER_MOD;
      pos += wsize;
      f += nused;
      carry_len = sm.carry_len; carry_dist = sm.carry_dist;
      __syncthreads();
    }
    __syncthreads();
  }
  if (pos != rec.out_len) return false;
  if (flushed + tid < pos) {               // the last partial 16-byte unit
    const uint32_t p = flushed + tid;
    const uint32_t d = sm.ring[p & RMASK];
    out[p] = (uint8_t)d;
    acc_a += d; acc_w += (unsigned long long)p * d;
  }
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
    ck = sm.crc;
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
