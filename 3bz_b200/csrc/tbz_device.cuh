// tbz_device.cuh — device-side shared definitions for the B200 inflate engine.
#pragma once
#include <stdint.h>
#include <cuda_runtime.h>
#include "threebz_cuda.h"

#define TBZ_FULL 0xffffffffu
#ifndef TBZ_EMU   // (tests/emu defines its own: the dynamic shared memory of the block)
#define TBZ_DYN_SMEM(name) extern __shared__ __align__(16) unsigned char name[]
#endif

// RFC 1951 tables in the reference's order: constants.lisp:36-61 (lengths at +32 there; split here)
__device__ __constant__ uint16_t c_len_base[32] = {3,4,5,6,7,8,9,10,11,13,15,17,19,23,27,31,35,43,51,59,67,83,99,
                                                   115,131,163,195,227,258,0,0,0};
__device__ __constant__ uint8_t c_len_extra[32] = {0,0,0,0,0,0,0,0,1,1,1,1,2,2,2,2,3,3,3,3,4,4,4,4,5,5,5,5,0,0,0,0};
__device__ __constant__ uint16_t c_dist_base[32] = {1,2,3,4,5,7,9,13,17,25,33,49,65,97,129,193,257,385,513,769,1025,
                                                    1537,2049,3073,4097,6145,8193,12289,16385,24577,0,0};
__device__ __constant__ uint8_t c_dist_extra[32] = {0,0,0,0,1,1,2,2,3,3,4,4,5,5,6,6,7,7,8,8,9,9,10,10,11,11,12,12,
                                                    13,13,0,0};
// constants.lisp:65-68
__device__ __constant__ uint8_t c_clen_order[19] = {16,17,18,0,8,7,9,6,10,5,11,4,12,3,13,2,14,1,15};
// x^(2^k) mod P for CRC-32 (reflected, P = 0xedb88320), k = 0..31: combine step of the
// segment-parallel CRC (checksums.lisp:196-210 computes the same function bytewise)
__device__ __constant__ uint32_t c_x2n[32] = {
  0x40000000u, 0x20000000u, 0x08000000u, 0x00800000u, 0x00008000u, 0xedb88320u, 0xb1e6b092u, 0xa06a2517u,
  0xed627daeu, 0x88d14467u, 0xd7bbfe6au, 0xec447f11u, 0x8e7ea170u, 0x6427800eu, 0x4d47bae0u, 0x09fe548fu,
  0x83852d0fu, 0x30362f1au, 0x7b5a9cc3u, 0x31fec169u, 0x9fec022au, 0x6c8dedc4u, 0x15d6874du, 0x5fde7a4eu,
  0xbad90e37u, 0x2e4e5eefu, 0x4eaba214u, 0xa8a472c0u, 0x429a969eu, 0x148d302au, 0xc40ba6d0u, 0xc4e22c3cu};

struct DMember {             // same layout as tbz_member
  const uint8_t *in; uint64_t in_len;
  uint8_t *out; uint64_t out_cap;
};

// ---------------------------------------------------------------------------------------------
// checksums (checksums.lisp restated as segment-parallel sums + combine)
// ---------------------------------------------------------------------------------------------
#define TBZ_ADLER_MOD 65521u
#define TBZ_CRC_POLY 0xedb88320u

__device__ __forceinline__ uint32_t crc_mulmod(uint32_t a, uint32_t b) {
  // a(x)*b(x) mod P in the reflected representation (bit 31 = x^0)
  uint32_t p = 0;
#pragma unroll 1
  for (int i = 0; i < 32; i++) {
    if (a & 0x80000000u) p ^= b;
    a <<= 1;
    b = (b >> 1) ^ ((b & 1) ? TBZ_CRC_POLY : 0);
    if (!a) break;
  }
  return p;
}
// x^(8*n) mod P
__device__ __forceinline__ uint32_t crc_x8n(uint64_t n) {
  uint32_t p = 0x80000000u;
  int k = 3;
  while (n) {
    if (n & 1) p = crc_mulmod(c_x2n[k & 31], p);
    n >>= 1; k++;
  }
  return p;
}
// crc(A||B) from crc(A), crc(B), len(B): finalized values in, finalized value out
__device__ __forceinline__ uint32_t crc_combine(uint32_t ca, uint32_t cb, uint64_t lenb) {
  return crc_mulmod(crc_x8n(lenb), ca) ^ cb;
}
__device__ __forceinline__ uint32_t crc_byte(uint32_t c, uint32_t b) {   // raw register update, bitwise
  c ^= b;
#pragma unroll
  for (int k = 0; k < 8; k++) c = (c >> 1) ^ ((c & 1) ? TBZ_CRC_POLY : 0);
  return c;
}
__device__ __forceinline__ void crc_table_init(uint32_t *tab /*256, shared*/, int tid, int nthreads) {
  for (int n = tid; n < 256; n += nthreads) tab[n] = crc_byte(0, (uint32_t)n);
}

// warp-parallel CRC-32 of p[0,n) (global or shared), table in shared memory.  Each lane takes one
// contiguous segment, the 32 segment CRCs are combined with x^(8 len) shifts.
__device__ inline uint32_t crc32_warp(const uint8_t *p, uint64_t n, const uint32_t *tab, int lane) {
  uint64_t seg = (n + 31) / 32;
  uint64_t lo = seg * lane, hi = lo + seg;
  if (lo > n) lo = n;
  if (hi > n) hi = n;
  uint32_t c = 0xffffffffu;
  for (uint64_t i = lo; i < hi; i++) c = (c >> 8) ^ tab[(c ^ p[i]) & 0xff];
  c ^= 0xffffffffu;
  if (lo == hi) c = 0;   // crc of the empty string
  // tree combine: at step s, lane l (multiple of 2s) absorbs lane l+s's segment
  uint64_t len = hi - lo;
#pragma unroll
  for (int s = 1; s < 32; s <<= 1) {
    uint32_t oc = __shfl_down_sync(TBZ_FULL, c, s);
    uint64_t ol = __shfl_down_sync(TBZ_FULL, len, s);
    if ((lane & (2 * s - 1)) == 0) {
      if (ol) c = crc_combine(c, oc, ol);
      len += ol;
    }
  }
  return __shfl_sync(TBZ_FULL, c, 0);
}

// warp-parallel Adler-32 of p[0,n): returns s1 | s2<<16 starting from (1,0).
// s1 = 1 + sum d_i ; s2 = n + sum (n-i) d_i   (mod 65521); tiles of 4096 bytes keep sums in 32 bits.
__device__ inline uint32_t adler32_warp(const uint8_t *p, uint64_t n, int lane, uint32_t s1 = 1, uint32_t s2 = 0) {
  for (uint64_t base = 0; base < n; base += 4096) {
    uint32_t m = (uint32_t)((n - base) < 4096 ? (n - base) : 4096);
    uint32_t a = 0, w = 0;   // a = sum d ; w = sum (m - j) d   over this lane's bytes
    for (uint32_t j = lane; j < m; j += 32) {
      uint32_t d = p[base + j];
      a += d; w += (m - j) * d;
    }
#pragma unroll
    for (int s = 16; s; s >>= 1) {
      a += __shfl_xor_sync(TBZ_FULL, a, s);
      w += __shfl_xor_sync(TBZ_FULL, w, s);     // <= 4096*4096*255/2 < 2^32
    }
    // appending a tile of m bytes: s2' = s2 + m*s1 + w ; s1' = s1 + a
    s2 = (uint32_t)(((uint64_t)s2 + (uint64_t)m * s1 + w) % TBZ_ADLER_MOD);
    s1 = (s1 + a) % TBZ_ADLER_MOD;
  }
  return s1 | (s2 << 16);
}
