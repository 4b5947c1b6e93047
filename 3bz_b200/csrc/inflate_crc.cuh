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
  uint32_t pw[32], pw2[32];                  // x^(8 seg (31 - l)) and x^(8 seg 32 (31 - w)): the weights of the two merge levels
  uint32_t xr;                               // x^(8 r): r = length of the member's last segment
  uint32_t seg_c, r_c;                       // the segment geometry the weights were computed for
  uint32_t part[NT / 32];
  uint32_t member;
};
constexpr size_t SMEM_BYTES = sizeof(Smem);

// CRC-32 of out[0, len) by the whole CTA; the value is returned to thread 0 (finalized: zlib's crc32()).
// Thread 0 takes the LAST segment (r bytes), threads 1 .. NT-1 the full segments in front of it, the last full one on
// thread NT-1: the state of thread u is then weighted with x^(8 (seg (NT-1-u) + r)), the same geometry for every span
// of the same length and alignment, so the weights are computed once per CTA and length (round 2 until here: every
// thread computed x^(8 seg) and ten squarings per member with the bit-serial multiplication — 52 % of the kernel's
// instructions, profiles/r2_crc_ncu.txt) and a span costs every thread ONE multiplication: segment state * weight,
// XOR over the warp, the 32 warp sums * weight, XOR, * x^(8 r).  States are raw register values (the first segment
// starts from 0xffffffff, the others from 0): linear in the data.  All threads must call; barriers inside.
__device__ inline uint32_t crc_span(const uint8_t *out, uint32_t len, Smem &sm, int tid) {
  const int lane = tid & 31, warp = tid >> 5;
#define TBZ_CRC_STEP(c, b) (c) = ((c) >> 8) ^ sm.tab[0][((c) ^ (b)) & 255u][lane]
#define TBZ_CRC_WORD(c, w) do { const uint32_t x_ = (c) ^ (w); \
    (c) = sm.tab[3][x_ & 255u][lane] ^ sm.tab[2][(x_ >> 8) & 255u][lane] ^ sm.tab[1][(x_ >> 16) & 255u][lane] ^ sm.tab[0][x_ >> 24][lane]; } while (0)
  // ---- geometry: segment t = [a_t, a_t+1), a_0 = 0, a_t = A0 + t seg (16-byte aligned addresses), clipped to len
  const uint32_t A0 = (uint32_t)((0u - (uintptr_t)out) & 15u);
  uint32_t seg = (len + NT - 1) / NT;
  seg = (seg + 63u) & ~63u;                 // whole 64-byte runs per thread
  if (!seg) seg = 64u;
  const uint32_t tlast = len > A0 ? (len - A0 - 1u) / seg : 0u;          // the last segment that is not empty
  const uint32_t r = len - (tlast ? A0 + tlast * seg : 0u);
  __syncthreads();                          // (the weights and sm.part of the span before this one have been read)
  if (sm.seg_c != seg || sm.r_c != r) {     // (uniform)
    __syncthreads();
    if (warp == 0) sm.pw[lane] = crc_x8n((uint64_t)seg * (31u - lane));
    else if (warp == 1) sm.pw2[lane] = crc_x8n((uint64_t)seg * 32u * (31u - lane));
    else if (tid == 64) { sm.xr = crc_x8n(r); sm.seg_c = seg; sm.r_c = r; }
    __syncthreads();
  }
  // ---- the thread's segment
  const int t = tid == 0 ? (int)tlast : (int)tlast - NT + tid;          // (tid >= 1: the full segments, the last one on NT-1)
  uint32_t lo = 0, hi = 0;
  if (t >= 0) {
    lo = t ? A0 + (uint32_t)t * seg : 0u;
    hi = A0 + ((uint32_t)t + 1u) * seg;
    if (hi > len) hi = len;
  }
  uint32_t c = t == 0 ? 0xffffffffu : 0u;
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
#undef TBZ_CRC_STEP
#undef TBZ_CRC_WORD
  // ---- merge: sum over u >= 1 of c_u x^(8 seg (NT-1-u)), times x^(8 r), plus the last segment's state
  uint32_t v = tid ? crc_mulmod(c, sm.pw[lane]) : 0u;
#pragma unroll
  for (int sft = 16; sft; sft >>= 1) v ^= __shfl_xor_sync(TBZ_FULL, v, sft);
  if (lane == 0) sm.part[warp] = v;
  __syncthreads();
  uint32_t crc = 0;
  if (warp == 0) {
    v = crc_mulmod(sm.part[lane], sm.pw2[lane]);
#pragma unroll
    for (int sft = 16; sft; sft >>= 1) v ^= __shfl_xor_sync(TBZ_FULL, v, sft);
    crc = (crc_mulmod(v, sm.xr) ^ c) ^ 0xffffffffu;     // (thread 0: c is the last segment's state)
  }
  return crc;
}

__device__ inline void crc_smem_init(Smem &sm, int tid) {
  if (tid < 256) {
    uint32_t e = crc_byte(0, (uint32_t)tid);
    for (int k = 0; k < 4; k++) {
#pragma unroll 8
      for (int l = 0; l < 32; l++) sm.tab[k][tid][l] = e;
      e = crc_byte(e, 0u);                       // one more zero byte behind it
    }
  }
  if (tid == 0) { sm.seg_c = 0; sm.r_c = 0; }
}

// One CTA per member (persistent): the CRC of its output against the trailer.
__global__ void __launch_bounds__(NT, 1)
k_member_crc(const DMember *members, tbz_result *results, uint32_t n, const tbzfast::P1Rec *recs, uint32_t *counters, uint32_t *todo) {
  TBZ_DYN_SMEM(smem_raw);
  Smem &sm = *reinterpret_cast<Smem *>(smem_raw);
  const int tid = threadIdx.x;
  crc_smem_init(sm, tid);
  for (;;) {
    __syncthreads();
    if (tid == 0) sm.member = atomicAdd(&counters[4], 1u);
    __syncthreads();
    const uint32_t i = sm.member;
    if (i >= n) break;
    if (recs[i].status != ST_CRC_PENDING) continue;
    const uint32_t crc = crc_span(members[i].out, (uint32_t)results[i].out_len, sm, tid);
    if (tid == 0) {
      const uint8_t *q = members[i].in + results[i].in_used - 8;      // CRC-32, ISIZE (the latter is not checked: gzip.lisp:99-106)
      const uint32_t tr = q[0] | ((uint32_t)q[1] << 8) | ((uint32_t)q[2] << 16) | ((uint32_t)q[3] << 24);
      if (tr == crc) results[i].checksum = crc;
      else todo[atomicAdd(&counters[1], 1u)] = i;                      // the sequential kernel reports the mismatch
    }
  }
}

// One large output (the split decode of a single member): parts[p] = CRC-32 of piece p (piece bytes each, the last one
// shorter); the host merges them with x^(8 n) shifts.
__global__ void __launch_bounds__(NT, 1)
k_span_crc(const uint8_t *out, uint64_t total, uint32_t piece, uint32_t *parts) {
  TBZ_DYN_SMEM(smem_raw);
  Smem &sm = *reinterpret_cast<Smem *>(smem_raw);
  const int tid = threadIdx.x;
  crc_smem_init(sm, tid);
  const uint64_t npieces = (total + piece - 1) / piece;
  for (uint64_t p = blockIdx.x; p < npieces; p += gridDim.x) {
    const uint64_t lo = p * piece;
    const uint32_t len = (uint32_t)(total - lo < piece ? total - lo : piece);
    const uint32_t crc = crc_span(out + lo, len, sm, tid);
    if (tid == 0) parts[p] = crc;
  }
}

}  // namespace tbzcrc
