// inflate_resolve.cuh — phase two of the batched fast path: LZ77 resolution of a token stream.
// One warp per member, 32 tokens per step (deflate.lisp:244-359 `copy-history`, restated):
//   * a warp prefix sum over the token lengths gives every token its output offset
//   * literals are stored at once
//   * matches are resolved in rounds: a match is ready when the bytes it reads lie below the
//     high-water mark (everything before the first still-pending token of the step); ready
//     matches copy concurrently, one byte per lane per iteration, overlapping matches read
//     through their period (i mod distance); the tail of matches longer than 32 bytes is copied
//     by the whole warp
//   * Adler-32 is folded in as the bytes are produced: s1 = 1 + sum d, s2 = N + N sum d - sum i d_i
//     is order independent, so every lane accumulates its own bytes; CRC-32 is a lane-parallel pass
//     over the finished member with x^(8 len) combines (checksums.lisp restated)
// The member's trailer is then checked exactly as zlib.lisp:80-96 / gzip.lisp:82-106 do; on any
// disagreement the member is queued for the sequential kernel, which owns the verdict rules.
#pragma once
#include "tbz_device.cuh"
#include "inflate_decode.cuh"

namespace tbzres {

using tbzfast::NO_SLAB;
using tbzfast::NT;
using tbzfast::P1Rec;
using tbzfast::SLAB_WORDS;
using tbzfast::SlabHdr;
using tbzfast::TOKCAP;
using tbzfast::TOK_EOB;
using tbzfast::TOK_MATCH;

struct Acc { unsigned long long a, w; };   // sum d ; sum i*d (reduced mod 65521 now and then)

__device__ __forceinline__ void acc_byte(Acc &c, uint32_t pos, uint32_t d) {
  c.a += d;
  c.w += (unsigned long long)pos * d;
}

__device__ inline bool resolve_member(const DMember &mem, int fmt, const P1Rec &rec, const uint32_t *__restrict__ slabs,
                                      tbz_result &res, const uint32_t *crc_tab, int lane) {
  uint8_t *out = mem.out;
  uint32_t pos = 0;
  Acc acc{0, 0};
  for (uint32_t s = rec.first_slab; s != NO_SLAB;) {
    const uint32_t *slab = slabs + (size_t)s * SLAB_WORDS;
    const SlabHdr *sh = reinterpret_cast<const SlabHdr *>(slab);
    const uint32_t *toks = slab + sizeof(SlabHdr) / 4;
    s = sh->next;
    for (int jb = 0; jb < NT; jb += 32) {
      const uint32_t gnv = sh->gn[jb + lane];
      uint32_t lanes = __ballot_sync(TBZ_FULL, gnv != 0);
      while (lanes) {
        const int jj = __ffs(lanes) - 1;
        lanes &= lanes - 1;
        const uint32_t gn = __shfl_sync(TBZ_FULL, gnv, jj);
        const uint32_t g = gn & 0xffffu, n = gn >> 16;
        const uint32_t *list = toks + (jb + jj) * TOKCAP;
        for (uint32_t k = g; k < n; k += 32) {
          // ---- one step: up to 32 consecutive tokens
          const uint32_t t = (k + lane < n) ? list[k + lane] : TOK_EOB;
          const bool is_m = (t & TOK_MATCH) != 0;
          const uint32_t len = is_m ? (t & 255u) + 3u : ((t & TOK_EOB) ? 0u : 1u);
          uint32_t incl = len;
#pragma unroll
          for (int sft = 1; sft < 32; sft <<= 1) {
            const uint32_t y = __shfl_up_sync(TBZ_FULL, incl, sft);
            if (lane >= sft) incl += y;
          }
          const uint32_t T = __shfl_sync(TBZ_FULL, incl, 31);
          const uint32_t dst = pos + incl - len;
          const uint32_t dist = ((t >> 8) & 0x7fffu) + 1u;
          if (__any_sync(TBZ_FULL, is_m && dist > dst)) return false;      // deflate.lisp:343-345
          if (!is_m && len) { out[dst] = (uint8_t)t; acc_byte(acc, dst, t & 255u); }
          const uint32_t src = dst - dist;
          const uint32_t need = dist < len ? dst : src + len;             // bytes [src, need) must be final
          bool pending = is_m;
          __syncwarp();
          uint32_t pm;
          while ((pm = __ballot_sync(TBZ_FULL, pending)) != 0) {
            const uint32_t hw = __shfl_sync(TBZ_FULL, dst, __ffs(pm) - 1);   // all output below is final
            const bool ready = pending && need <= hw;
            const uint32_t n1 = ready ? (len < 32u ? len : 32u) : 0u;
            const uint32_t nmax = __reduce_max_sync(TBZ_FULL, n1);
            if (dist >= len) {
              for (uint32_t i = 0; i < nmax; i++)
                if (i < n1) { const uint32_t d = out[src + i]; out[dst + i] = (uint8_t)d; acc_byte(acc, dst + i, d); }
            } else {
              for (uint32_t i = 0; i < nmax; i++)
                if (i < n1) { const uint32_t d = out[src + i % dist]; out[dst + i] = (uint8_t)d; acc_byte(acc, dst + i, d); }
            }
            uint32_t lm = __ballot_sync(TBZ_FULL, ready && len > 32u);
            while (lm) {                                                    // long matches: whole warp
              const int l = __ffs(lm) - 1;
              lm &= lm - 1;
              const uint32_t bdst = __shfl_sync(TBZ_FULL, dst, l), bsrc = __shfl_sync(TBZ_FULL, src, l);
              const uint32_t blen = __shfl_sync(TBZ_FULL, len, l), bdist = __shfl_sync(TBZ_FULL, dist, l);
              for (uint32_t i = 32 + lane; i < blen; i += 32) {
                const uint32_t d = out[bsrc + (bdist >= blen ? i : i % bdist)];
                out[bdst + i] = (uint8_t)d;
                acc_byte(acc, bdst + i, d);
              }
            }
            pending = pending && !ready;
            __syncwarp();                                                   // stores visible to the next round
          }
          pos += T;
          if (acc.w >> 62) acc.w %= TBZ_ADLER_MOD;
        }
      }
    }
  }
  if (pos != rec.out_len) return false;
  __syncwarp();
  // ---- checksum of the whole member
  uint32_t ck = 0;
  if (fmt == TBZ_ZLIB) {
    unsigned long long a = acc.a, w = acc.w % TBZ_ADLER_MOD;
#pragma unroll
    for (int sft = 16; sft; sft >>= 1) { a += __shfl_xor_sync(TBZ_FULL, a, sft); w += __shfl_xor_sync(TBZ_FULL, w, sft); }
    const unsigned long long N = pos % TBZ_ADLER_MOD, S = a % TBZ_ADLER_MOD;
    const uint32_t s1 = (uint32_t)((1 + S) % TBZ_ADLER_MOD);
    const uint32_t s2 = (uint32_t)((N + N * S + (unsigned long long)TBZ_ADLER_MOD * 32 - w % TBZ_ADLER_MOD) % TBZ_ADLER_MOD);
    ck = s1 | (s2 << 16);
  } else if (fmt == TBZ_GZIP) {
    ck = crc32_warp(out, pos, crc_tab, lane);
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

}  // namespace tbzres
