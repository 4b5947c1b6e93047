// kernels.cuh — the __global__ entry points of the batched path (launched from runtime.cu).  They live in
// a header so that the CPU test suite can run the same source under tests/emu (a SIMT emulator; test
// infrastructure only — the product library is this code compiled by nvcc for sm_100a, nothing else).
#pragma once
#include "tbz_device.cuh"
#include "inflate_seq.cuh"
#include "inflate_decode.cuh"
#include "huff_decode.cuh"
#include "inflate_copy.cuh"
#include "inflate_crc.cuh"

// =============================================================================================
// kernels
// =============================================================================================
#define SEQ_WARPS 4

__global__ void __launch_bounds__(SEQ_WARPS * 32)
k_inflate_seq(const DMember *members, tbz_result *results, uint32_t n, int fmt,
              const uint32_t *todo, const uint32_t *todo_count) {
  __shared__ tbzseq::WarpSmem sm[SEQ_WARPS];
  __shared__ uint32_t crc_tab[256];
  crc_table_init(crc_tab, threadIdx.x, blockDim.x);
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint32_t i = blockIdx.x * SEQ_WARPS + warp;
  if (todo_count) n = *todo_count;
  if (i >= n) return;
  if (todo) i = todo[i];
  tbzseq::inflate_member(members[i], fmt, results[i], sm[warp], crc_tab, lane);
}

// A session's member (sessions: decompress / replace-output-buffer with chunked input), one warp: resumes at the
// last block boundary an earlier call reached and leaves the new one behind (tbzseq::Resume, device-resident).
__global__ void __launch_bounds__(32)
k_inflate_session(DMember member, int fmt, tbz_result *result, tbzseq::Resume *rs) {
  __shared__ tbzseq::WarpSmem sm;
  __shared__ uint32_t crc_tab[256];
  crc_table_init(crc_tab, threadIdx.x, blockDim.x);
  __syncthreads();
  tbzseq::inflate_member(member, fmt, *result, sm, crc_tab, threadIdx.x & 31, rs);
}

// counters: [0] next member for phase one, [1] members queued for the sequential kernel,
//           [2] slabs handed out, [3] next member for phase two, [4] next member for the gzip CRC kernel
// Phase one (huff_decode.cuh), persistent CTAs of WPC independent warps: each warp pulls the next member from a
// global counter and decodes it into token slabs; members it cannot prove clean are queued for k_inflate_seq.
// scratch: tbzhd::SCRATCH_BYTES per warp of the grid.
__global__ void __launch_bounds__(tbzhd::NT, TBZ_HD_MINBLOCKS)
k_inflate_decode(const DMember *members, uint32_t n, int fmt, tbzfast::P1Rec *recs,
                 uint32_t *slabs, uint32_t nslabs, uint32_t *counters, uint32_t *todo, unsigned char *scratch) {
  TBZ_DYN_SMEM(smem_raw);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  tbzhd::WSmem &sm = reinterpret_cast<tbzhd::WSmem *>(smem_raw)[warp];
  uint32_t *const gck = reinterpret_cast<uint32_t *>(scratch + ((size_t)blockIdx.x * tbzhd::WPC + warp) * tbzhd::SCRATCH_BYTES);
  for (;;) {
    uint32_t i = 0;
    if (lane == 0) i = atomicAdd(&counters[0], 1u);
    i = __shfl_sync(TBZ_FULL, i, 0);
    if (i >= n) break;
    const bool ok = tbzhd::decode_member(members[i], fmt, recs[i], sm, gck, slabs, nslabs, &counters[2], lane);
    __syncwarp();
    if (!ok && lane == 0) { recs[i].status = 0; todo[atomicAdd(&counters[1], 1u)] = i; }
  }
}

// Phase two, persistent CTAs: one CTA per member resolves the token stream into bytes through a
// shared-memory window and checks the trailer (inflate_copy.cuh).  The alternatives that were built and
// measured — byte-parallel rank queries and lock-step lanes (round 1), one warp per member with 64-bit or
// 16-byte tokens (round 2: inflate_decode2/resolve2, huff_decode/lz_resolve) — live under experiments/ with
// their numbers in profiles/ and DESIGN.md.
namespace tbzp2 = tbzcp;
#ifndef TBZ_P2_MINBLOCKS
#define TBZ_P2_MINBLOCKS 3
#endif
__global__ void __launch_bounds__(tbzp2::NT, TBZ_P2_MINBLOCKS)
k_inflate_resolve(const DMember *members, tbz_result *results, uint32_t n, int fmt,
                  const tbzfast::P1Rec *recs, const uint32_t *slabs, uint32_t *counters, uint32_t *todo) {
  TBZ_DYN_SMEM(smem_raw);
  tbzp2::Smem &sm = *reinterpret_cast<tbzp2::Smem *>(smem_raw);
  const int tid = threadIdx.x;
  if (fmt == TBZ_GZIP && !tbzp2::CRC_SEPARATE) {
    crc_table_init(sm.crc_tab, tid, tbzp2::NT);
    for (uint32_t k = tid; k < tbzp2::WB / 16 + 4; k += tbzp2::NT) sm.x16[k] = crc_x8n(16ull * k);
  }
  for (;;) {
    __syncthreads();
    if (tid == 0) sm.member = atomicAdd(&counters[3], 1u);
    __syncthreads();
    const uint32_t i = sm.member;
    if (i >= n) break;
    if (!recs[i].status) continue;
    const bool ok = tbzp2::resolve_member(members[i], fmt, recs[i], slabs, results[i], sm, tid);
    if (!ok && tid == 0) todo[atomicAdd(&counters[1], 1u)] = i;
    if (ok && tid == 0 && tbzp2::CRC_SEPARATE && fmt == TBZ_GZIP) const_cast<tbzfast::P1Rec *>(recs)[i].status = tbzcrc::ST_CRC_PENDING;
  }
}
