// kernels.cuh — the __global__ entry points of the batched path (launched from runtime.cu).  They live in
// a header so that the CPU test suite can run the same source under tests/emu (a SIMT emulator; test
// infrastructure only — the product library is this code compiled by nvcc for sm_100a, nothing else).
#pragma once
#include "tbz_device.cuh"
#include "inflate_seq.cuh"
#include "inflate_decode.cuh"
#include "huff_decode.cuh"
#include "lz_resolve.cuh"
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

// counters: [0] next member for phase one, [1] members queued for the sequential kernel,
//           [2] 16-byte units of the token heap handed out, [3] next member for phase two, [4] next member for the gzip CRC kernel
// Phase one (huff_decode.cuh), persistent CTAs of WPC independent warps: each warp pulls the next member from a
// global counter and decodes it into blocks of the token heap (through its own scratch lists: SCRATCH_BYTES per warp
// of the grid); members it cannot prove clean are queued for k_inflate_seq.
__global__ void __launch_bounds__(tbzhd::NT, TBZ_HD_MINBLOCKS)
k_inflate_decode(const DMember *members, uint32_t n, int fmt, tbzfast::P1Rec *recs, unsigned char *scratch_all,
                 uint4 *heap, uint32_t heap_units, uint32_t *counters, uint32_t *todo) {
  TBZ_DYN_SMEM(smem_raw);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  tbzhd::WSmem &sm = reinterpret_cast<tbzhd::WSmem *>(smem_raw)[warp];
  uint32_t *const scratch = reinterpret_cast<uint32_t *>(scratch_all + ((size_t)blockIdx.x * tbzhd::WPC + warp) * tbzhd::SCRATCH_BYTES);
  for (;;) {
    uint32_t i = 0;
    if (lane == 0) i = atomicAdd(&counters[0], 1u);
    i = __shfl_sync(TBZ_FULL, i, 0);
    if (i >= n) break;
    const bool ok = tbzhd::decode_member(members[i], fmt, recs[i], sm, scratch, heap, heap_units, &counters[2], lane);
    __syncwarp();
    if (!ok && lane == 0) { recs[i].status = 0; todo[atomicAdd(&counters[1], 1u)] = i; }
  }
}

// Phase two (lz_resolve.cuh), persistent CTAs of WPC independent warps: ONE WARP per member, a history ring per warp,
// no CTA barrier.  (The round-1 design — one CTA per member, windows, pointer jumping — and the
// first round-2 pair live under experiments/ with their numbers in profiles/.)
__global__ void __launch_bounds__(tbzlz::NT, TBZ_LZ_MINBLOCKS)
k_inflate_resolve(const DMember *members, tbz_result *results, uint32_t n, int fmt,
                  const tbzfast::P1Rec *recs, const uint4 *heap, uint32_t *counters, uint32_t *todo) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t ring = tbzlz::smem_base() + tbzlz::PAD + (uint32_t)warp * tbzlz::H;   // shared-space address of this warp's ring
  for (;;) {
    uint32_t i = 0;
    if (lane == 0) i = atomicAdd(&counters[3], 1u);
    i = __shfl_sync(TBZ_FULL, i, 0);
    if (i >= n) break;
    if (!recs[i].status) continue;
    const bool ok = tbzlz::resolve_member(members[i], fmt, recs[i], heap, results[i], ring, lane);
    __syncwarp();
    if (!ok && lane == 0) todo[atomicAdd(&counters[1], 1u)] = i;
    if (ok && lane == 0 && fmt == TBZ_GZIP) const_cast<tbzfast::P1Rec *>(recs)[i].status = tbzcrc::ST_CRC_PENDING;
  }
}
