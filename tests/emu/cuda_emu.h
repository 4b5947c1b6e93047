// cuda_emu.h — a small SIMT emulator for the CPU test suite.  TEST INFRASTRUCTURE ONLY: it lets
// `pytest -m "not gpu"` execute the *same* device source (3bz_b200/csrc/*.cuh) that nvcc compiles
// for sm_100a, so that kernel logic (indexing, warp collectives, barriers, edge cases) is checked
// against the oracle without a GPU.  It is not a fallback: nothing under 3bz_b200/ includes it and
// the product library has no CPU path.
//
// Model: one CUDA thread = one ucontext fiber; one block runs on one OS thread, its fibers scheduled
// round-robin; a fiber runs until it reaches a warp collective or a block barrier it cannot complete,
// then yields.  Collectives check that every lane named in the mask arrives with the same operation
// (an exited or diverged lane aborts the run: on hardware that is undefined behaviour).  Misaligned
// vector accesses are caught by building with -fsanitize=alignment.
#pragma once
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <ucontext.h>
#include <atomic>
#include <functional>
#include <thread>
#include <vector>

#define TBZ_EMU 1
#define __device__
#define __host__
#define __global__
#define __forceinline__ inline __attribute__((always_inline))
#define __constant__ const
#define __launch_bounds__(...)
#define __noinline__ __attribute__((noinline))
#define __align__(n) alignas(n)
#define __shared__ static thread_local

struct uint2 { uint32_t x, y; };
struct alignas(16) uint4 { uint32_t x, y, z, w; };
struct alignas(8) uint2a { uint32_t x, y; };
static inline uint2 make_uint2(uint32_t x, uint32_t y) { return uint2{x, y}; }
static inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { return uint4{x, y, z, w}; }
struct dim3 { unsigned x, y, z; dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {} };

namespace emu {

enum Op { OP_NONE = 0, OP_SYNCWARP, OP_SHFL, OP_SHFL_UP, OP_SHFL_DOWN, OP_SHFL_XOR, OP_BALLOT, OP_ANY, OP_ALL, OP_MATCH };

struct Warp {
  uint32_t arrived = 0, alive = 0, mask = 0, gen = 0;
  int op = OP_NONE;
  uint64_t in[32], aux[32], out[2][32];
};

struct Fiber {
  ucontext_t ctx;
  std::vector<char> stack;
  bool done = false;
};

struct Block {
  unsigned nthreads = 0;
  std::vector<Fiber> fibers;
  std::vector<Warp> warps;
  ucontext_t sched;
  unsigned cur = 0;
  // block barrier
  unsigned bar_arrived = 0, bar_alive = 0, bar_gen = 0;
  int bar_or = 0, bar_or_out[2] = {0, 0};
  uint64_t epoch = 0;                 // bumps whenever anything completes (deadlock detection)
  std::vector<unsigned char> dyn;
  std::function<void()> body;
  dim3 bidx, bdim, gdim;
};

extern thread_local Block *g_blk;
void die(const char *what);
inline void yield() { Block *b = g_blk; swapcontext(&b->fibers[b->cur].ctx, &b->sched); }
unsigned char *dyn_smem();

// Generic warp collective: deposit (val, aux); the last lane of `mask` to arrive computes every lane's result.
template <class F>
inline uint64_t collective(int op, uint32_t mask, uint64_t val, uint64_t aux, F compute) {
  Block *b = g_blk;
  const unsigned tid = b->cur, lane = tid & 31;
  Warp &w = b->warps[tid >> 5];
  if (!(mask >> lane & 1)) die("collective: calling lane not in mask");
  if (mask & ~w.alive) die("collective: mask names a lane that has exited (undefined behaviour on hardware)");
  if (w.arrived == 0) { w.op = op; w.mask = mask; }
  else if (w.op != op || w.mask != mask) die("collective: lanes of one warp arrived at different collectives / masks (diverged warp)");
  w.in[lane] = val; w.aux[lane] = aux;
  w.arrived |= 1u << lane;
  const uint32_t g = w.gen;
  if (w.arrived == mask) {
    compute(w, w.out[g & 1]);
    w.arrived = 0; w.gen = g + 1; b->epoch++;
  } else {
    while (w.gen == g) yield();
  }
  return w.out[g & 1][lane];
}

void block_barrier(int pred, int *or_out);
void launch_impl(dim3 grid, dim3 block, size_t smem, const std::function<void()> &body, int os_threads);

}  // namespace emu

struct EmuIdx { unsigned x, y, z; };
extern thread_local EmuIdx threadIdx, blockIdx, blockDim, gridDim;

// ---- warp collectives ---------------------------------------------------------------------------
static inline void __syncwarp(uint32_t mask = 0xffffffffu) {
  emu::collective(emu::OP_SYNCWARP, mask, 0, 0, [](emu::Warp &, uint64_t *) {});
}
template <class T> static inline T emu_bits_to(uint64_t v) { T t; memcpy(&t, &v, sizeof(T)); return t; }
template <class T> static inline uint64_t emu_to_bits(T t) { uint64_t v = 0; memcpy(&v, &t, sizeof(T)); return v; }

template <class T> static inline T __shfl_sync(uint32_t mask, T v, int src, int width = 32) {
  (void)width;
  return emu_bits_to<T>(emu::collective(emu::OP_SHFL, mask, emu_to_bits(v), (uint64_t)(src & 31), [](emu::Warp &w, uint64_t *out) {
    for (int l = 0; l < 32; l++) if (w.mask >> l & 1) { int s = (int)w.aux[l]; out[l] = (w.mask >> s & 1) ? w.in[s] : w.in[l]; }
  }));
}
template <class T> static inline T __shfl_up_sync(uint32_t mask, T v, unsigned d, int width = 32) {
  (void)width;
  return emu_bits_to<T>(emu::collective(emu::OP_SHFL_UP, mask, emu_to_bits(v), d, [](emu::Warp &w, uint64_t *out) {
    for (int l = 0; l < 32; l++) if (w.mask >> l & 1) { int s = l - (int)w.aux[l]; out[l] = (s >= 0 && (w.mask >> s & 1)) ? w.in[s] : w.in[l]; }
  }));
}
template <class T> static inline T __shfl_down_sync(uint32_t mask, T v, unsigned d, int width = 32) {
  (void)width;
  return emu_bits_to<T>(emu::collective(emu::OP_SHFL_DOWN, mask, emu_to_bits(v), d, [](emu::Warp &w, uint64_t *out) {
    for (int l = 0; l < 32; l++) if (w.mask >> l & 1) { int s = l + (int)w.aux[l]; out[l] = (s < 32 && (w.mask >> s & 1)) ? w.in[s] : w.in[l]; }
  }));
}
template <class T> static inline T __shfl_xor_sync(uint32_t mask, T v, int x, int width = 32) {
  (void)width;
  return emu_bits_to<T>(emu::collective(emu::OP_SHFL_XOR, mask, emu_to_bits(v), (uint64_t)x, [](emu::Warp &w, uint64_t *out) {
    for (int l = 0; l < 32; l++) if (w.mask >> l & 1) { int s = l ^ (int)w.aux[l]; out[l] = (s < 32 && (w.mask >> s & 1)) ? w.in[s] : w.in[l]; }
  }));
}
static inline uint32_t __ballot_sync(uint32_t mask, int pred) {
  return (uint32_t)emu::collective(emu::OP_BALLOT, mask, pred ? 1 : 0, 0, [](emu::Warp &w, uint64_t *out) {
    uint32_t r = 0;
    for (int l = 0; l < 32; l++) if ((w.mask >> l & 1) && w.in[l]) r |= 1u << l;
    for (int l = 0; l < 32; l++) out[l] = r;
  });
}
static inline int __any_sync(uint32_t mask, int pred) { return __ballot_sync(mask, pred) != 0; }
static inline int __all_sync(uint32_t mask, int pred) { return __ballot_sync(mask, pred) == mask; }
template <class T> static inline uint32_t __match_any_sync(uint32_t mask, T v) {
  return (uint32_t)emu::collective(emu::OP_MATCH, mask, emu_to_bits(v), 0, [](emu::Warp &w, uint64_t *out) {
    for (int l = 0; l < 32; l++) if (w.mask >> l & 1) {
      uint32_t r = 0;
      for (int k = 0; k < 32; k++) if ((w.mask >> k & 1) && w.in[k] == w.in[l]) r |= 1u << k;
      out[l] = r;
    }
  });
}
static inline uint32_t __reduce_max_sync(uint32_t mask, uint32_t v) {
  return (uint32_t)emu::collective(emu::OP_BALLOT + 100, mask, v, 0, [](emu::Warp &w, uint64_t *out) {
    uint64_t r = 0;
    for (int l = 0; l < 32; l++) if ((w.mask >> l & 1) && w.in[l] > r) r = w.in[l];
    for (int l = 0; l < 32; l++) out[l] = r;
  });
}
static inline uint32_t __activemask() { return 0xffffffffu; }

// ---- block barriers ------------------------------------------------------------------------------
static inline void __syncthreads() { emu::block_barrier(0, nullptr); }
static inline int __syncthreads_or(int pred) { int r = 0; emu::block_barrier(pred, &r); return r; }
static inline void __threadfence_block() {}
static inline void __threadfence() {}

// ---- scalar intrinsics ---------------------------------------------------------------------------
static inline uint32_t __funnelshift_r(uint32_t lo, uint32_t hi, uint32_t s) { return (uint32_t)((((uint64_t)hi << 32) | lo) >> (s & 31)); }
static inline uint32_t __funnelshift_l(uint32_t lo, uint32_t hi, uint32_t s) { return (uint32_t)(((((uint64_t)hi << 32) | lo) << (s & 31)) >> 32); }
static inline uint32_t __brev(uint32_t x) { uint32_t r = 0; for (int i = 0; i < 32; i++) r |= ((x >> i) & 1u) << (31 - i); return r; }
static inline int __popc(uint32_t x) { return __builtin_popcount(x); }
static inline int __popcll(unsigned long long x) { return __builtin_popcountll(x); }
static inline int __clz(int x) { return x ? __builtin_clz((unsigned)x) : 32; }
static inline int __ffs(int x) { return __builtin_ffs(x); }
static inline int __ffsll(long long x) { return __builtin_ffsll(x); }
static inline uint32_t __dp4a(uint32_t a, uint32_t b, uint32_t c) {
  for (int i = 0; i < 4; i++) c += ((a >> (8 * i)) & 255u) * ((b >> (8 * i)) & 255u);
  return c;
}
static inline uint32_t __byte_perm(uint32_t x, uint32_t y, uint32_t s) {
  const uint64_t v = ((uint64_t)y << 32) | x;
  uint32_t r = 0;
  for (int i = 0; i < 4; i++) {
    const uint32_t sel = (s >> (4 * i)) & 15u;
    uint32_t byte = (uint32_t)(v >> (8 * (sel & 7))) & 255u;
    if (sel & 8) byte = (byte & 0x80) ? 0xff : 0;
    r |= byte << (8 * i);
  }
  return r;
}
template <class T> static inline T __ldg(const T *p) { return *p; }
template <class T> static inline T __ldcg(const T *p) { return *p; }
template <class T> static inline T __ldcs(const T *p) { return *p; }
static inline long long clock64() { return 0; }
template <class T> static inline T __ldca(const T *p) { return *p; }
static inline uint32_t __umulhi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }

static inline uint32_t min(uint32_t a, uint32_t b) { return a < b ? a : b; }
static inline uint32_t max(uint32_t a, uint32_t b) { return a > b ? a : b; }
static inline int min(int a, int b) { return a < b ? a : b; }
static inline int max(int a, int b) { return a > b ? a : b; }
static inline unsigned long long min(unsigned long long a, unsigned long long b) { return a < b ? a : b; }
static inline unsigned long long max(unsigned long long a, unsigned long long b) { return a > b ? a : b; }
static inline unsigned long min(unsigned long a, unsigned long b) { return a < b ? a : b; }
static inline unsigned long max(unsigned long a, unsigned long b) { return a > b ? a : b; }

// ---- atomics (blocks run on several OS threads) ---------------------------------------------------
template <class T> static inline T atomicAdd(T *p, T v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
template <class T> static inline T atomicOr(T *p, T v) { return __atomic_fetch_or(p, v, __ATOMIC_RELAXED); }
template <class T> static inline T atomicAnd(T *p, T v) { return __atomic_fetch_and(p, v, __ATOMIC_RELAXED); }
template <class T> static inline T atomicExch(T *p, T v) { return __atomic_exchange_n(p, v, __ATOMIC_RELAXED); }
template <class T> static inline T atomicMin(T *p, T v) {
  T o = __atomic_load_n(p, __ATOMIC_RELAXED);
  while (v < o && !__atomic_compare_exchange_n(p, &o, v, true, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
  return o;
}
template <class T> static inline T atomicMax(T *p, T v) {
  T o = __atomic_load_n(p, __ATOMIC_RELAXED);
  while (v > o && !__atomic_compare_exchange_n(p, &o, v, true, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
  return o;
}
template <class T> static inline T atomicCAS(T *p, T cmp, T v) {
  __atomic_compare_exchange_n(p, &cmp, v, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED);
  return cmp;
}

// ---- launch ---------------------------------------------------------------------------------------
// emu_launch(kernel, grid, block, dynamic_smem_bytes, args...): every block on one of `os_threads` OS threads.
extern int emu_os_threads;
template <class K, class... A>
static inline void emu_launch(K kernel, dim3 grid, dim3 block, size_t smem, A... args) {
  emu::launch_impl(grid, block, smem, [=]() { kernel(args...); }, emu_os_threads);
}
#define TBZ_DYN_SMEM(name) unsigned char *name = ::emu::dyn_smem()
