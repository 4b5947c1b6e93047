// emu_driver.cpp — runs the batched kernel sequence of runtime.cu::launch_kernels under the SIMT
// emulator (tests/emu/cuda_emu.h).  TEST INFRASTRUCTURE ONLY: built and loaded by tests/test_emu.py.
#include <cuda_runtime.h>
#include <cstring>
#include <vector>
#include "kernels.cuh"
#include "inflate_split.cuh"

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
  const uint64_t round_bytes = (uint64_t)tbzfast::NL * tbzfast::S_MAX / 8;
  uint64_t want = 0;
  for (uint64_t i = 0; i < n; i++) want += 2 * (m[i].in_len / round_bytes) + 4;
  const uint32_t nslabs = (uint32_t)want;
  std::vector<uint32_t> slabs((size_t)nslabs * tbzfast::SLAB_WORDS);
  (void)variant;
  {
    const unsigned dec_grid = std::min<unsigned>((nn + tbzhd::WPC - 1) / tbzhd::WPC, 16);
    std::vector<unsigned char> scratch((size_t)dec_grid * tbzhd::WPC * tbzhd::SCRATCH_BYTES + 16);
    emu_launch(k_inflate_decode, dim3(dec_grid), dim3(tbzhd::NT), sizeof(tbzhd::WSmem) * tbzhd::WPC,
               (const DMember *)dm.data(), nn, fmt, recs.data(), slabs.data(), nslabs, counters.data(), todo.data(), scratch.data());
    emu_launch(k_inflate_resolve, dim3(std::min<unsigned>(nn, 16)), dim3(tbzp2::NT), sizeof(tbzp2::Smem),
               (const DMember *)dm.data(), r, nn, fmt, (const tbzfast::P1Rec *)recs.data(), (const uint32_t *)slabs.data(),
               counters.data(), todo.data());
  }
  if (fmt == TBZ_GZIP)
    emu_launch(tbzcrc::k_member_crc, dim3(std::min<unsigned>(nn, 4)), dim3(tbzcrc::NT), tbzcrc::SMEM_BYTES,
               (const DMember *)dm.data(), r, nn, (const tbzfast::P1Rec *)recs.data(), counters.data(), todo.data());
  if (!(variant & 2))          // (variant bit 1: leave the members the fast kernels gave up on as they are, for debugging)
    emu_launch(k_inflate_seq, dim3((nn + SEQ_WARPS - 1) / SEQ_WARPS), dim3(SEQ_WARPS * 32), 0,
               (const DMember *)dm.data(), r, nn, fmt, (const uint32_t *)todo.data(), (const uint32_t *)counters.data() + 1);
  return (int)counters[1];     // members that took the sequential kernel
}

// The block-start search of the split decode (tbzfast::find_block_start), one warp: first dynamic-block start in bits
// [from, to) of the stream `data` (4-byte aligned copy, `nbytes` long); 0xffffffff = none.
__global__ void k_emu_find(const uint32_t *words, uint32_t nbits, uint32_t from, uint32_t to, uint32_t *result) {
  TBZ_DYN_SMEM(smem_raw);
  tbzfast::WSmem &sm = *reinterpret_cast<tbzfast::WSmem *>(smem_raw);
  tbzfast::In in;
  in.w = words; in.pos0 = 0; in.end = nbits; in.nwords = (nbits + 31) >> 5;
  const uint32_t r = tbzfast::find_block_start(in, from, to, sm, threadIdx.x & 31);
  if (threadIdx.x == 0) *result = r;
}

extern "C" uint32_t emu_find_block_start(const uint8_t *data, uint64_t nbytes, uint32_t from, uint32_t to) {
  emu_os_threads = 1;
  std::vector<uint32_t> words((nbytes + 3) / 4 + 2, 0);
  memcpy(words.data(), data, nbytes);
  uint32_t result = 0;
  emu_launch(k_emu_find, dim3(1), dim3(32), sizeof(tbzfast::WSmem), (const uint32_t *)words.data(), (uint32_t)(nbytes * 8), from, to, &result);
  return result;
}

// ---- kernels of the split decode of one large member that do not need a whole member to be tested ----------------
// Adler-32 of data[0, n) from k_split_adler's per-segment sums, chained as runtime.cu does
extern "C" uint32_t emu_split_adler(const uint8_t *data, uint64_t n) {
  emu_os_threads = 1;
  const uint64_t nseg = (n + tbzsplit::ASEG - 1) / tbzsplit::ASEG;
  std::vector<uint32_t> parts(2 * std::max<uint64_t>(1, nseg), 0);
  if (nseg) emu_launch(tbzsplit::k_split_adler, dim3((unsigned)((nseg + 7) / 8)), dim3(256), 0, data, n, parts.data());
  uint64_t s1 = 1, s2 = 0;
  for (uint64_t s = 0; s < nseg; s++) {
    const uint64_t len = std::min<uint64_t>(tbzsplit::ASEG, n - s * tbzsplit::ASEG);
    s2 = (s2 + len * s1 + parts[2 * s + 1]) % TBZ_ADLER_MOD;
    s1 = (s1 + parts[2 * s]) % TBZ_ADLER_MOD;
  }
  return (uint32_t)(s1 | (s2 << 16));
}
// out[v] = in[v] o in[v - stride] over nv maps of 32 Ki entries (k_tail_compose)
extern "C" void emu_tail_compose(const uint16_t *in, uint16_t *out, uint32_t nv, uint32_t stride) {
  emu_os_threads = 1;
  emu_launch(tbzsplit::k_tail_compose, dim3(nv, tbzsplit::TAILW / 2048), dim3(256), 0, in, out, stride);
}
// symbols -> bytes (k_split_translate); offs[nchunks + 1], out holds the final tails already
extern "C" void emu_split_translate(const uint64_t *offs, uint32_t nchunks, const uint16_t *sym, uint8_t *out, uint64_t total, uint32_t grid) {
  emu_os_threads = 1;
  emu_launch(tbzsplit::k_split_translate, dim3(grid), dim3(256), 0, offs, nchunks, sym, out, total);
}
