// emu_driver.cpp — runs the batched kernel sequence of runtime.cu::launch_kernels under the SIMT
// emulator (tests/emu/cuda_emu.h).  TEST INFRASTRUCTURE ONLY: built and loaded by tests/test_emu.py.
#include <cuda_runtime.h>
#include <vector>
#include "kernels.cuh"

extern "C" int emu_inflate_batch(int fmt, const tbz_member *m, uint64_t n, tbz_result *r, uint32_t flags, int os_threads, int variant) {
  emu_os_threads = os_threads > 0 ? os_threads : 1;
  if (!n) return 0;
  std::vector<DMember> dm(n);
  for (uint64_t i = 0; i < n; i++) dm[i] = DMember{m[i].in, m[i].in_len, m[i].out, m[i].out_cap};
  std::vector<uint32_t> counters(64, 0), todo(n, 0);
  const uint32_t nn = (uint32_t)n;
  if (flags & TBZ_FLAG_NO_FASTPATH) {
    emu_launch(k_inflate_seq, dim3((nn + SEQ_WARPS - 1) / SEQ_WARPS), dim3(SEQ_WARPS * 32), 0,
               (const DMember *)dm.data(), r, nn, fmt, (const uint32_t *)nullptr, (const uint32_t *)nullptr);
    return 0;
  }
  std::vector<tbzfast::P1Rec> recs(n);
  // the token heap: 16-byte units; ~1 token byte per output byte on text, more for short matches
  uint64_t units = 64;
  for (uint64_t i = 0; i < n; i++) units += (6 * m[i].out_cap + 8 * m[i].in_len) / 16 + 64;
  const uint32_t heap_units = (uint32_t)units;
  std::vector<uint4> heap(heap_units);
  (void)variant;
  const unsigned dec_grid = std::min<unsigned>((nn + tbzhd::WPC - 1) / tbzhd::WPC, 16);
  std::vector<uint4> scratch((size_t)dec_grid * tbzhd::WPC * tbzhd::SCRATCH_BYTES / 16);
  emu_launch(k_inflate_decode, dim3(dec_grid), dim3(tbzhd::NT), sizeof(tbzhd::WSmem) * tbzhd::WPC,
             (const DMember *)dm.data(), nn, fmt, recs.data(), (unsigned char *)scratch.data(), heap.data(), heap_units, counters.data(), todo.data());
  emu_launch(k_inflate_resolve, dim3(std::min<unsigned>((nn + tbzlz::WPC - 1) / tbzlz::WPC, 16)), dim3(tbzlz::NT), tbzlz::SMEM_BYTES,
             (const DMember *)dm.data(), r, nn, fmt, (const tbzfast::P1Rec *)recs.data(), (const uint4 *)heap.data(),
             counters.data(), todo.data());
  if (fmt == TBZ_GZIP)
    emu_launch(tbzcrc::k_member_crc, dim3(std::min<unsigned>(nn, 16)), dim3(tbzcrc::NT), 0,
               (const DMember *)dm.data(), r, nn, (const tbzfast::P1Rec *)recs.data(), counters.data(), todo.data());
  if (!(variant & 2))          // (variant bit 1: leave the members the fast kernels gave up on as they are, for debugging)
    emu_launch(k_inflate_seq, dim3((nn + SEQ_WARPS - 1) / SEQ_WARPS), dim3(SEQ_WARPS * 32), 0,
               (const DMember *)dm.data(), r, nn, fmt, (const uint32_t *)todo.data(), (const uint32_t *)counters.data() + 1);
  return (int)counters[1];     // members that took the sequential kernel
}
