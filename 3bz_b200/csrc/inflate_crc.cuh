// inflate_crc.cuh — CRC-32 of the decompressed members of a gzip batch, as a kernel of its own
// (checksums.lisp:177-210 `crc32/table`, restated segment-parallel with x^(8n) combines), and the
// trailer check of gzip.lisp:82-106.
//
// Inside phase two the CRC costs a table walk per 16-byte unit plus a 32-step GF(2) multiplication
// (no carry-less multiply on sm_100a) between barriers: +45 % on 1 MiB gzip members.  As a streaming
// pass over the finished output it is table steps in long per-thread runs and one multiplication per
// thread and member.  Round 1 walked one byte per step (one dependent shared-memory lookup per byte:
// 1.2 - 1.6 TB/s, the warps waiting on each other's lookups); this is slicing-by-4: four bytes per step,
// four INDEPENDENT lookups in four tables (T0 = the byte table, Tk[b] = the CRC of byte b followed by k
// zero bytes), 4 instructions per byte instead of 7.  Every table is replicated per lane
// (tab[k][index][lane]: every lane reads its own bank, no conflicts): 128 KB of shared memory, one CTA
// of 1024 threads per SM.  The output is read with 16-byte loads; it is re-read once (L2 / HBM).
// A member whose CRC disagrees with its trailer is queued for the sequential kernel, which owns the verdict.
#pragma once
#include "tbz_device.cuh"
#include "inflate_decode.cuh"

namespace tbzcrc {

constexpr int NT = 1024;
constexpr uint32_t ST_CRC_PENDING = 3;      // P1Rec.status: resolved by phase two, CRC and trailer compare still to do

struct Smem {
  uint32_t tab[4][256][32];                  // the four slicing tables, one copy per lane (bank)
  uint32_t part[NT], plen[NT];
  uint32_t member;
};
constexpr size_t SMEM_BYTES = sizeof(Smem);

__global__ void __launch_bounds__(NT, 1)
k_member_crc(const DMember *members, tbz_result *results, uint32_t n, const tbzfast::P1Rec *recs, uint32_t *counters, uint32_t *todo) {
  TBZ_DYN_SMEM(smem_raw);
  Smem &sm = *reinterpret_cast<Smem *>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31;
  if (tid < 256) {
    uint32_t e = crc_byte(0, (uint32_t)tid);
    for (int k = 0; k < 4; k++) {
#pragma unroll 8
      for (int l = 0; l < 32; l++) sm.tab[k][tid][l] = e;
      e = crc_byte(e, 0u);                       // one more zero byte behind it
    }
  }
#define TBZ_CRC_STEP(c, b) (c) = ((c) >> 8) ^ sm.tab[0][((c) ^ (b)) & 255u][lane]
#define TBZ_CRC_WORD(c, w) do { const uint32_t x_ = (c) ^ (w); \
    (c) = sm.tab[3][x_ & 255u][lane] ^ sm.tab[2][(x_ >> 8) & 255u][lane] ^ sm.tab[1][(x_ >> 16) & 255u][lane] ^ sm.tab[0][x_ >> 24][lane]; } while (0)
  for (;;) {
    __syncthreads();
    if (tid == 0) sm.member = atomicAdd(&counters[4], 1u);
    __syncthreads();
    const uint32_t i = sm.member;
    if (i >= n) break;
    if (recs[i].status != ST_CRC_PENDING) continue;
    const uint8_t *out = members[i].out;
    const uint32_t len = (uint32_t)results[i].out_len;
    uint32_t seg = (len + NT - 1) / NT;
    seg = (seg + 63u) & ~63u;                 // whole 64-byte runs per thread
    uint32_t lo = seg * tid < len ? seg * tid : len, hi = lo + seg < len ? lo + seg : len;
    const uint32_t mylen = hi - lo;
    uint32_t c = 0xffffffffu;
    while (lo < hi && ((uintptr_t)(out + lo) & 15u)) { TBZ_CRC_STEP(c, out[lo]); lo++; }
    for (; lo + 64u <= hi; lo += 64u) {         // four loads back to back: both halves of a 32-byte sector are asked for
      uint4 v4[4];                               // before anything can evict it
#pragma unroll
      for (int k = 0; k < 4; k++) v4[k] = *reinterpret_cast<const uint4 *>(out + lo + 16u * k);
#pragma unroll
      for (int k = 0; k < 4; k++) {
        TBZ_CRC_WORD(c, v4[k].x); TBZ_CRC_WORD(c, v4[k].y); TBZ_CRC_WORD(c, v4[k].z); TBZ_CRC_WORD(c, v4[k].w);
      }
    }
    for (; lo + 16u <= hi; lo += 16u) {
      const uint4 v = *reinterpret_cast<const uint4 *>(out + lo);
      TBZ_CRC_WORD(c, v.x); TBZ_CRC_WORD(c, v.y); TBZ_CRC_WORD(c, v.z); TBZ_CRC_WORD(c, v.w);
    }
    for (; lo < hi; lo++) TBZ_CRC_STEP(c, out[lo]);
    c ^= 0xffffffffu;
    if (!mylen) c = 0;
    // pairwise merge: crc(A || B) = crc(A) * x^(8 |B|) + crc(B); the shift of a full segment is squared per level
    uint32_t plen = mylen;
    uint32_t shift = crc_x8n(seg);
    for (int s = 1; s < NT; s <<= 1) {
      sm.part[tid] = c; sm.plen[tid] = plen;
      __syncthreads();
      if ((tid & (2 * s - 1)) == 0 && tid + s < NT) {
        const uint32_t oc = sm.part[tid + s], ol = sm.plen[tid + s];
        if (ol) {
          const uint32_t f = (ol == seg * (uint32_t)s) ? shift : crc_x8n(ol);
          c = crc_mulmod(f, c) ^ oc;
          plen += ol;
        }
      }
      shift = crc_mulmod(shift, shift);
      __syncthreads();
    }
    if (tid == 0) {
      const uint8_t *q = members[i].in + results[i].in_used - 8;      // CRC-32, ISIZE (the latter is not checked: gzip.lisp:99-106)
      const uint32_t t = q[0] | ((uint32_t)q[1] << 8) | ((uint32_t)q[2] << 16) | ((uint32_t)q[3] << 24);
      if (t == c) results[i].checksum = c;
      else todo[atomicAdd(&counters[1], 1u)] = i;                      // the sequential kernel reports the mismatch
    }
  }
#undef TBZ_CRC_STEP
#undef TBZ_CRC_WORD
}

}  // namespace tbzcrc
