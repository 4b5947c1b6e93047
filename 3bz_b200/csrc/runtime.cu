// runtime.cu — host runtime of libthreebz_cuda.so: contexts, memory, batches, sessions,
// multi-GPU partitioning.  The only translation unit; kernels live in the .cuh files.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <chrono>
#include <atomic>
#include <functional>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "threebz_cuda.h"
#include "tbz_device.cuh"
#include "inflate_seq.cuh"
#include "inflate_decode.cuh"
#include "inflate_resolve.cuh"
#include "inflate_copy.cuh"
#include "inflate_crc.cuh"
#include "kernels.cuh"
#include "inflate_split.cuh"

// =============================================================================================
// host objects
// =============================================================================================
static const uint64_t kSlabPoolBytes = 24ull << 30;  // upper bound of the token slab pool per batch (B200: 180 GB)
static const uint64_t kSplitMinBytes = 4ull << 20;   // members at least this large are split across the GPU
static const uint64_t kSplitChunkBytes =      // compressed bytes per chunk of a split member (TBZ_SPLIT_CHUNK_KB: tuning)
    (getenv("TBZ_SPLIT_CHUNK_KB") ? std::max<uint64_t>(16, strtoull(getenv("TBZ_SPLIT_CHUNK_KB"), nullptr, 10)) : 224ull) << 10;

struct DevBlock { void *p; size_t size; bool used; };

struct tbz_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;          // the stream new work is enqueued on (main_stream, or a pipeline stream)
  static const int kPipeStreams = 8;
  cudaStream_t main_stream = nullptr, pstream[kPipeStreams] = {};
  cudaEvent_t ev0 = nullptr, ev1 = nullptr, tev0 = nullptr, tev1 = nullptr;
  cudaEvent_t kev[4] = {nullptr, nullptr, nullptr, nullptr};   // TBZ_KTIME=1: events between the kernels of a launch
  bool ktime = false, ktime_quiet = false;
  float last_kms[3] = {0.f, 0.f, 0.f};
  std::vector<DevBlock> pool;
  void *stage_in = nullptr;  size_t stage_in_cap = 0;    // pinned staging
  void *stage_out = nullptr; size_t stage_out_cap = 0;
  void *stage_res = nullptr; size_t stage_res_cap = 0;   // pinned landing zone for result records (the caller's array may be pageable)
  std::string last_error;
  uint64_t launches = 0;
  int sm_count = 0;
};

static thread_local std::string g_last_error;

static int32_t fail(tbz_ctx *ctx, int32_t code, const char *what, cudaError_t e = cudaSuccess) {
  char buf[512];
  if (e != cudaSuccess) snprintf(buf, sizeof buf, "%s: %s", what, cudaGetErrorString(e));
  else snprintf(buf, sizeof buf, "%s", what);
  if (ctx) ctx->last_error = buf;
  g_last_error = buf;
  return code;
}
#define CK(ctx, call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return fail((ctx), TBZ_E_CUDA, #call, e_); } while (0)

static int32_t dev_alloc(tbz_ctx *ctx, size_t n, void **p) {
  if (n == 0) n = 256;
  n = (n + 255) & ~(size_t)255;
  int best = -1;
  for (size_t i = 0; i < ctx->pool.size(); i++) {
    DevBlock &b = ctx->pool[i];
    if (!b.used && b.size >= n && (best < 0 || b.size < ctx->pool[best].size)) best = (int)i;
  }
  if (best >= 0 && ctx->pool[best].size <= 2 * n + (1 << 20)) {
    ctx->pool[best].used = true; *p = ctx->pool[best].p; return TBZ_OK;
  }
  void *q = nullptr;
  cudaError_t e = cudaMalloc(&q, n);
  if (e != cudaSuccess) {
    // release cached blocks and retry once
    for (auto it = ctx->pool.begin(); it != ctx->pool.end();)
      if (!it->used) { cudaFree(it->p); it = ctx->pool.erase(it); } else ++it;
    e = cudaMalloc(&q, n);
    if (e != cudaSuccess) { cudaGetLastError(); return fail(ctx, TBZ_E_NOMEM, "cudaMalloc", e); }
  }
  ctx->pool.push_back({q, n, true});
  *p = q;
  return TBZ_OK;
}
static void dev_release(tbz_ctx *ctx, void *p) {
  if (!p) return;
  for (auto &b : ctx->pool) if (b.p == p) { b.used = false; return; }
}

static int32_t ensure_stage(tbz_ctx *ctx, void **p, size_t *cap, size_t n) {
  if (*cap >= n) return TBZ_OK;
  if (*p) cudaFreeHost(*p);
  *p = nullptr; *cap = 0;
  size_t want = n + n / 4 + 4096;
  cudaError_t e = cudaHostAlloc(p, want, cudaHostAllocDefault);
  if (e != cudaSuccess) { cudaGetLastError(); return fail(ctx, TBZ_E_NOMEM, "cudaHostAlloc", e); }
  *cap = want;
  return TBZ_OK;
}

struct tbz_batch {
  tbz_ctx *ctx = nullptr;
  int format = 0;
  uint32_t flags = 0;
  uint64_t n = 0;
  bool device_ptrs = false;
  std::vector<tbz_member> host;        // caller's members (host mode)
  std::vector<uint64_t> in_off, out_off;
  uint64_t in_total = 0, out_total = 0;
  bool in_direct = false;              // caller's inputs are one dense span: DMA straight from it
  bool out_direct = false;             // caller's outputs are exactly adjacent: DMA straight into them
  bool copies_issued = false;          // (pipelined path) the D2H copies of the produced bytes are already in the stream
  const uint8_t *in_span = nullptr; uint8_t *out_span = nullptr;
  void *d_in = nullptr, *d_out = nullptr, *d_members = nullptr, *d_results = nullptr;
  void *d_slabs = nullptr, *d_counters = nullptr, *d_todo = nullptr, *d_recs = nullptr, *d_scratch = nullptr;
  int fast_grid = 0, res_grid = 0;
  uint32_t nslabs = 0;
  bool launched = false;
  cudaStream_t stream = nullptr;       // the stream this batch lives on
  cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev_res = nullptr;   // kernels start / end; (pipelined path) results in pinned memory
  tbz_result *eager_res = nullptr;     // pipelined sub-batch: results and outputs are copied back right behind the kernels
  std::vector<DMember> dm;             // device view of the members (kept alive for the asynchronous upload)
  std::vector<std::pair<uint64_t, DMember>> big;   // members decoded by the split path (index, device view)
};

// =============================================================================================
// library / context
// =============================================================================================
extern "C" int32_t tbz_abi_version(void) { return TBZ_ABI_VERSION; }

extern "C" int32_t tbz_device_count(int32_t *n) {
  if (!n) return TBZ_E_ARG;
  int c = 0;
  cudaError_t e = cudaGetDeviceCount(&c);
  if (e != cudaSuccess) { cudaGetLastError(); *n = 0; return fail(nullptr, TBZ_E_NO_DEVICE, "cudaGetDeviceCount", e); }
  *n = c;
  return TBZ_OK;
}

extern "C" int32_t tbz_ctx_create(int32_t device, uint64_t flags, tbz_ctx **out) {
  (void)flags;
  if (!out) return TBZ_E_ARG;
  *out = nullptr;
  int c = 0;
  cudaError_t e = cudaGetDeviceCount(&c);
  if (e != cudaSuccess || c == 0) { cudaGetLastError(); return fail(nullptr, TBZ_E_NO_DEVICE, "no CUDA device (this engine has no CPU fallback)", e); }
  if (device < 0 || device >= c) return fail(nullptr, TBZ_E_ARG, "device index out of range");
  tbz_ctx *ctx = new tbz_ctx();
  ctx->device = device;
  CK(ctx, cudaSetDevice(device));
  CK(ctx, cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
  ctx->main_stream = ctx->stream;
  for (auto &ps : ctx->pstream) CK(ctx, cudaStreamCreateWithFlags(&ps, cudaStreamNonBlocking));
  CK(ctx, cudaEventCreate(&ctx->ev0)); CK(ctx, cudaEventCreate(&ctx->ev1));
  CK(ctx, cudaEventCreate(&ctx->tev0)); CK(ctx, cudaEventCreate(&ctx->tev1));
  CK(ctx, cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, device));
  if (const char *kt = getenv("TBZ_KTIME")) ctx->ktime = kt[0] == '1';
  if (ctx->ktime) for (auto &e : ctx->kev) CK(ctx, cudaEventCreate(&e));
  *out = ctx;
  return TBZ_OK;
}

extern "C" int32_t tbz_ctx_destroy(tbz_ctx *ctx) {
  if (!ctx) return TBZ_OK;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  for (auto &b : ctx->pool) cudaFree(b.p);
  if (ctx->stage_in) cudaFreeHost(ctx->stage_in);
  if (ctx->stage_out) cudaFreeHost(ctx->stage_out);
  if (ctx->stage_res) cudaFreeHost(ctx->stage_res);
  cudaEventDestroy(ctx->ev0); cudaEventDestroy(ctx->ev1);
  cudaEventDestroy(ctx->tev0); cudaEventDestroy(ctx->tev1);
  cudaStreamDestroy(ctx->main_stream);
  for (auto &ps : ctx->pstream) cudaStreamDestroy(ps);
  delete ctx;
  return TBZ_OK;
}

extern "C" const char *tbz_strerror(int32_t s) {
  switch (s) {
    case TBZ_OK: return "ok";
    case TBZ_E_CUDA: return "CUDA runtime error";
    case TBZ_E_NO_DEVICE: return "no CUDA device (no CPU fallback exists)";
    case TBZ_E_ARG: return "bad argument";
    case TBZ_E_NOMEM: return "out of memory";
    case TBZ_E_BUFFER_SWITCH: return "can't switch buffers without filling old one yet.";
    case TBZ_E_STATE: return "session is not usable in this state";
    default: return "unknown status";
  }
}

extern "C" const char *tbz_verdict_name(int32_t v) {
  switch (v) {
    case TBZ_FINISHED: return "finished";
    case TBZ_INPUT_UNDERRUN: return "input-underrun";
    case TBZ_OUTPUT_OVERFLOW: return "output-overflow";
    case TBZ_ERR_BLOCK_TYPE: return "reserved block type";
    case TBZ_ERR_STORED_LEN: return "stored block LEN/NLEN mismatch";
    case TBZ_ERR_OVERSUBSCRIBED: return "too many entries in huffman table";
    case TBZ_ERR_INCOMPLETE: return "incomplete huffman table";
    case TBZ_ERR_REPEAT_NO_PREV: return "tried to repeat length without previous length";
    case TBZ_ERR_REPEAT_OVERRUN: return "code length repeat runs past table";
    case TBZ_ERR_INVALID_SYMBOL: return "invalid huffman code";
    case TBZ_ERR_DISTANCE_TOO_FAR: return "distance reaches before start of output";
    case TBZ_ERR_ZLIB_FCHECK: return "invalid zlib header checksum";
    case TBZ_ERR_ZLIB_METHOD: return "invalid zlib compression type";
    case TBZ_ERR_ZLIB_WINDOW: return "invalid window size in zlib header";
    case TBZ_ERR_ZLIB_DICT: return "preset dictionary not supported yet";
    case TBZ_ERR_GZIP_MAGIC: return "bad gzip magic";
    case TBZ_ERR_GZIP_METHOD: return "unknown compression method";
    case TBZ_ERR_GZIP_RESERVED: return "reserved flag bits set";
    case TBZ_ERR_GZIP_HCRC: return "gzip header crc mismatch";
    case TBZ_ERR_CHECKSUM: return "checksum mismatch";
    case TBZ_ERR_TREE_TOO_LARGE: return "huffman table does not fit";
    default: return "unknown verdict";
  }
}

extern "C" const char *tbz_ctx_last_error(tbz_ctx *ctx) {
  return ctx ? ctx->last_error.c_str() : g_last_error.c_str();
}
extern "C" int32_t tbz_ctx_synchronize(tbz_ctx *ctx) {
  if (!ctx) return TBZ_E_ARG;
  CK(ctx, cudaStreamSynchronize(ctx->stream));
  return TBZ_OK;
}
extern "C" int32_t tbz_ctx_stream(tbz_ctx *ctx, void **s) {
  if (!ctx || !s) return TBZ_E_ARG;
  *s = (void *)ctx->stream;
  return TBZ_OK;
}
extern "C" int32_t tbz_ctx_timer_start(tbz_ctx *ctx) {
  if (!ctx) return TBZ_E_ARG;
  CK(ctx, cudaEventRecord(ctx->tev0, ctx->stream));
  return TBZ_OK;
}
extern "C" int32_t tbz_ctx_timer_stop(tbz_ctx *ctx, float *ms) {
  if (!ctx || !ms) return TBZ_E_ARG;
  CK(ctx, cudaEventRecord(ctx->tev1, ctx->stream));
  CK(ctx, cudaEventSynchronize(ctx->tev1));
  CK(ctx, cudaEventElapsedTime(ms, ctx->tev0, ctx->tev1));
  return TBZ_OK;
}
extern "C" int32_t tbz_ctx_kernel_timing(tbz_ctx *ctx, int32_t enable) {
  if (!ctx) return TBZ_E_ARG;
  CK(ctx, cudaSetDevice(ctx->device));
  if (enable && !ctx->kev[0]) for (auto &e : ctx->kev) CK(ctx, cudaEventCreate(&e));
  if (enable) { if (!ctx->ktime) ctx->ktime_quiet = true; ctx->ktime = true; }
  else if (ctx->ktime_quiet) { ctx->ktime = false; ctx->ktime_quiet = false; }
  return TBZ_OK;
}
extern "C" int32_t tbz_ctx_last_kernel_ms(tbz_ctx *ctx, float *ms3) {
  if (!ctx || !ms3) return TBZ_E_ARG;
  for (int i = 0; i < 3; i++) ms3[i] = ctx->last_kms[i];
  return TBZ_OK;
}
extern "C" int32_t tbz_ctx_launch_count(tbz_ctx *ctx, uint64_t *n) {
  if (!ctx || !n) return TBZ_E_ARG;
  *n = ctx->launches;
  return TBZ_OK;
}

// =============================================================================================
// memory
// =============================================================================================
extern "C" int32_t tbz_host_alloc(uint64_t n, void **p) {
  if (!p) return TBZ_E_ARG;
  cudaError_t e = cudaHostAlloc(p, n ? n : 1, cudaHostAllocPortable);
  if (e != cudaSuccess) { cudaGetLastError(); *p = nullptr; return fail(nullptr, e == cudaErrorMemoryAllocation ? TBZ_E_NOMEM : TBZ_E_NO_DEVICE, "cudaHostAlloc", e); }
  return TBZ_OK;
}
extern "C" int32_t tbz_host_free(void *p) {
  if (p) cudaFreeHost(p);
  return TBZ_OK;
}
extern "C" int32_t tbz_host_register(void *p, uint64_t n, uint32_t flags) {
  (void)flags;
  if (!p || !n) return TBZ_E_ARG;
  cudaError_t e = cudaHostRegister(p, n, cudaHostRegisterPortable | cudaHostRegisterReadOnly);
  if (e != cudaSuccess) { cudaGetLastError(); e = cudaHostRegister(p, n, cudaHostRegisterPortable); }
  if (e != cudaSuccess) { cudaGetLastError(); return fail(nullptr, TBZ_E_CUDA, "cudaHostRegister", e); }
  return TBZ_OK;
}
extern "C" int32_t tbz_host_unregister(void *p) {
  if (!p) return TBZ_E_ARG;
  cudaError_t e = cudaHostUnregister(p);
  if (e != cudaSuccess) { cudaGetLastError(); return fail(nullptr, TBZ_E_CUDA, "cudaHostUnregister", e); }
  return TBZ_OK;
}
extern "C" int32_t tbz_device_alloc(tbz_ctx *ctx, uint64_t n, void **p) {
  if (!ctx || !p) return TBZ_E_ARG;
  CK(ctx, cudaSetDevice(ctx->device));
  cudaError_t e = cudaMalloc(p, n ? n : 1);
  if (e != cudaSuccess) { cudaGetLastError(); return fail(ctx, TBZ_E_NOMEM, "cudaMalloc", e); }
  return TBZ_OK;
}
extern "C" int32_t tbz_device_free(tbz_ctx *ctx, void *p) {
  if (!ctx) return TBZ_E_ARG;
  CK(ctx, cudaSetDevice(ctx->device));
  if (p) CK(ctx, cudaFree(p));
  return TBZ_OK;
}
extern "C" int32_t tbz_memcpy_h2d(tbz_ctx *ctx, void *dst, const void *src, uint64_t n) {
  if (!ctx) return TBZ_E_ARG;
  CK(ctx, cudaSetDevice(ctx->device));
  CK(ctx, cudaMemcpyAsync(dst, src, n, cudaMemcpyHostToDevice, ctx->stream));
  CK(ctx, cudaStreamSynchronize(ctx->stream));
  return TBZ_OK;
}
extern "C" int32_t tbz_memcpy_d2h(tbz_ctx *ctx, void *dst, const void *src, uint64_t n) {
  if (!ctx) return TBZ_E_ARG;
  CK(ctx, cudaSetDevice(ctx->device));
  CK(ctx, cudaMemcpyAsync(dst, src, n, cudaMemcpyDeviceToHost, ctx->stream));
  CK(ctx, cudaStreamSynchronize(ctx->stream));
  return TBZ_OK;
}
extern "C" void tbz_free(void *p) { free(p); }

// =============================================================================================
// batches
// =============================================================================================
static void parallel_for(size_t n, size_t bytes, const std::function<void(size_t, size_t)> &f) {
  unsigned hw = std::thread::hardware_concurrency();
  size_t nt = bytes < (8u << 20) ? 1 : std::min<size_t>(hw ? hw : 1, 16);
  if (nt <= 1 || n < 2 * nt) { f(0, n); return; }
  std::vector<std::thread> th;
  size_t per = (n + nt - 1) / nt;
  for (size_t t = 0; t < nt; t++) {
    size_t lo = t * per, hi = std::min(n, lo + per);
    if (lo >= hi) break;
    th.emplace_back([=, &f] { f(lo, hi); });
  }
  for (auto &t : th) t.join();
}

extern "C" int32_t tbz_batch_destroy(tbz_batch *b) {
  if (!b) return TBZ_OK;
  tbz_ctx *ctx = b->ctx;
  cudaSetDevice(ctx->device);
  if (b->launched || b->n) cudaStreamSynchronize(b->stream);
  if (b->ev0) cudaEventDestroy(b->ev0);
  if (b->ev1) cudaEventDestroy(b->ev1);
  if (b->ev_res) cudaEventDestroy(b->ev_res);
  dev_release(ctx, b->d_in); dev_release(ctx, b->d_out);
  dev_release(ctx, b->d_members); dev_release(ctx, b->d_results);
  dev_release(ctx, b->d_slabs); dev_release(ctx, b->d_counters); dev_release(ctx, b->d_todo); dev_release(ctx, b->d_recs);
  dev_release(ctx, b->d_scratch);
  delete b;
  return TBZ_OK;
}

extern "C" int32_t tbz_batch_prepare(tbz_ctx *ctx, int32_t format, const tbz_member *m, uint64_t n,
                                     uint32_t flags, tbz_batch **out) {
  if (!ctx || !out || (n && !m) || format < 0 || format > 2 || n > 0x7fffffffull) return fail(ctx, TBZ_E_ARG, "tbz_batch_prepare: bad argument");
  *out = nullptr;
  CK(ctx, cudaSetDevice(ctx->device));
  tbz_batch *b = new tbz_batch();
  b->ctx = ctx; b->format = format; b->flags = flags; b->n = n;
  b->stream = ctx->stream;
  CK(ctx, cudaEventCreate(&b->ev0)); CK(ctx, cudaEventCreate(&b->ev1));
  CK(ctx, cudaEventCreateWithFlags(&b->ev_res, cudaEventDisableTiming));
  b->device_ptrs = (flags & TBZ_FLAG_DEVICE_PTRS) != 0;
  int32_t rc;
#define PCK(x) do { rc = (x); if (rc != TBZ_OK) { tbz_batch_destroy(b); return rc; } } while (0)
  PCK(dev_alloc(ctx, std::max<uint64_t>(1, n) * sizeof(DMember), &b->d_members));
  PCK(dev_alloc(ctx, std::max<uint64_t>(1, n) * sizeof(tbz_result), &b->d_results));
  if (!(flags & TBZ_FLAG_NO_FASTPATH) && n) {
    int occ = 0;
    cudaFuncSetAttribute(k_inflate_decode, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(tbzhd::WSmem) * tbzhd::WPC));
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_inflate_decode, tbzhd::NT, sizeof(tbzhd::WSmem) * tbzhd::WPC);
    b->fast_grid = (int)std::min<uint64_t>((n + tbzhd::WPC - 1) / tbzhd::WPC, (uint64_t)ctx->sm_count * std::max(1, occ));
    cudaFuncSetAttribute(k_inflate_resolve, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(tbzp2::Smem));
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_inflate_resolve, tbzp2::NT, sizeof(tbzp2::Smem));
    b->res_grid = (int)std::min<uint64_t>(n, (uint64_t)ctx->sm_count * std::max(1, occ));
    // token slabs: one per round; a round covers at most ~31 KiB of compressed input and never
    // crosses a block boundary.  Members that find the pool empty go to the sequential kernel.
    uint64_t want = 0;
    // a round covers up to NL * S_MAX bits; blocks end rounds early and a lane that fills its token
    // list shortens them, hence the factor 2 and the slack
    const uint64_t round_bytes = (uint64_t)tbzfast::NL * tbzfast::S_MAX / 8;
    for (uint64_t i = 0; i < n; i++) want += 2 * (m[i].in_len / round_bytes) + 4;
    const uint64_t slab_bytes = (uint64_t)tbzfast::SLAB_WORDS * 4;
    const uint64_t cap = std::max<uint64_t>(64, kSlabPoolBytes / slab_bytes);
    b->nslabs = (uint32_t)std::min<uint64_t>(want, cap);
    PCK(dev_alloc(ctx, (size_t)b->nslabs * slab_bytes, &b->d_slabs));
    PCK(dev_alloc(ctx, (size_t)b->fast_grid * tbzhd::WPC * tbzhd::SCRATCH_BYTES, &b->d_scratch));
    PCK(dev_alloc(ctx, 256, &b->d_counters));
    PCK(dev_alloc(ctx, n * 4, &b->d_todo));
    PCK(dev_alloc(ctx, n * sizeof(tbzfast::P1Rec), &b->d_recs));
  }
  std::vector<DMember> &dm = b->dm;
  dm.resize(n);
  if (b->device_ptrs) {
    for (uint64_t i = 0; i < n; i++) dm[i] = DMember{m[i].in, m[i].in_len, m[i].out, m[i].out_cap};
  } else {
    b->host.assign(m, m + n);
    b->in_off.resize(n); b->out_off.resize(n);
    // dense, ordered inputs -> one DMA straight from the caller's span
    bool in_dense = n > 0, out_adj = n > 0;
    uint64_t in_sum = 0;
    for (uint64_t i = 0; i < n; i++) {
      in_sum += m[i].in_len;
      if (i + 1 < n) {
        // (one span only when nothing but alignment padding lies between the members: a gap may be memory the
        // caller never handed over — another allocation, an unmapped page)
        if (m[i].in + m[i].in_len > m[i + 1].in || (uint64_t)(m[i + 1].in - (m[i].in + m[i].in_len)) >= 16) in_dense = false;
        if (m[i].out + m[i].out_cap != m[i + 1].out) out_adj = false;
      }
    }
    if (in_dense) {
      uint64_t span = (uint64_t)((m[n - 1].in + m[n - 1].in_len) - m[0].in);
      {
        b->in_direct = true; b->in_span = m[0].in; b->in_total = span;
        for (uint64_t i = 0; i < n; i++) b->in_off[i] = (uint64_t)(m[i].in - m[0].in);
      }
    }
    if (!b->in_direct) {
      uint64_t o = 0;
      for (uint64_t i = 0; i < n; i++) { b->in_off[i] = o; o += (m[i].in_len + 15) & ~15ull; }
      b->in_total = o;
    }
    if (out_adj) {
      b->out_direct = true; b->out_span = m[0].out;
      uint64_t o = 0;
      for (uint64_t i = 0; i < n; i++) { b->out_off[i] = o; o += m[i].out_cap; }
      b->out_total = o;
    } else {
      uint64_t o = 0;
      for (uint64_t i = 0; i < n; i++) { b->out_off[i] = o; o += (m[i].out_cap + 15) & ~15ull; }
      b->out_total = o;
    }
    PCK(dev_alloc(ctx, b->in_total + 16, &b->d_in));
    PCK(dev_alloc(ctx, b->out_total + 16, &b->d_out));
    for (uint64_t i = 0; i < n; i++)
      dm[i] = DMember{(const uint8_t *)b->d_in + b->in_off[i], m[i].in_len,
                      (uint8_t *)b->d_out + b->out_off[i], m[i].out_cap};
  }
  if (!(flags & (TBZ_FLAG_NO_SPLIT | TBZ_FLAG_NO_FASTPATH)))
    for (uint64_t i = 0; i < n; i++)
      if (dm[i].in_len >= kSplitMinBytes) {           // the batched kernels see an empty member; split_inflate fills the result
        b->big.push_back({i, dm[i]});
        dm[i].in_len = 0; dm[i].out_cap = 0;
      }
  if (n) {
    cudaError_t e = cudaMemcpyAsync(b->d_members, dm.data(), n * sizeof(DMember), cudaMemcpyHostToDevice, b->stream);
    if (e != cudaSuccess) { tbz_batch_destroy(b); return fail(ctx, TBZ_E_CUDA, "upload member table", e); }
  }
#undef PCK
  *out = b;
  return TBZ_OK;
}

// ---------------------------------------------------------------------------------------------
// one large member, split across the whole GPU (inflate_split.cuh).  Synchronous.  Returns TBZ_OK
// with *handled = false when the member has to take the ordinary path (odd header, broken chain,
// truncated or corrupt stream, too small an output buffer): the sequential kernel then owns the verdict.
// ---------------------------------------------------------------------------------------------
static uint32_t crc_mulmod_h(uint32_t a, uint32_t b) {
  uint32_t p = 0;
  for (int i = 0; i < 32 && a; i++) {
    if (a & 0x80000000u) p ^= b;
    a <<= 1;
    b = (b >> 1) ^ ((b & 1) ? 0xedb88320u : 0);
  }
  return p;
}
static uint32_t crc_x8n_h(uint64_t n) {            // x^(8n) mod P
  uint32_t p = 0x80000000u, sq = 0x00800000u;     // x^0, x^8
  while (n) {
    if (n & 1) p = crc_mulmod_h(sq, p);
    sq = crc_mulmod_h(sq, sq);
    n >>= 1;
  }
  return p;
}

static int32_t split_inflate(tbz_ctx *ctx, int fmt, const DMember &m, tbz_result *d_result, bool *handled) {
  using tbzsplit::Chunk;
  *handled = false;
  cudaStream_t st = ctx->stream;
  auto t_prev = std::chrono::steady_clock::now();
  auto stage = [&](const char *name) {                // TBZ_KTIME=1: wall time per stage (each ends in a stream sync)
    if (!ctx->ktime || ctx->ktime_quiet) return;
    cudaStreamSynchronize(st);
    auto t = std::chrono::steady_clock::now();
    fprintf(stderr, "[tbz split] %-10s %8.3f ms\n", name, std::chrono::duration<double, std::milli>(t - t_prev).count());
    t_prev = t;
  };
  // ---- wrapper header, on the host (zlib.lisp:108-126; gzip.lisp:113-260 with FEXTRA / FNAME / FCOMMENT / FHCRC:
  // `gzip file` always writes a name).  Anything but a clean header leaves the member to the sequential kernel,
  // which owns the verdicts.
  uint64_t hdr_bytes = 0;
  if (fmt == TBZ_ZLIB) {
    uint8_t head[2] = {0, 0};
    CK(ctx, cudaMemcpyAsync(head, m.in, 2, cudaMemcpyDeviceToHost, st));
    CK(ctx, cudaStreamSynchronize(st));
    if ((head[0] * 256 + head[1]) % 31 || (head[0] & 15) != 8 || (head[0] >> 4) > 7 || (head[1] & 32)) return TBZ_OK;
    hdr_bytes = 2;
  } else if (fmt == TBZ_GZIP) {
    std::vector<uint8_t> head;
    for (uint64_t take = 4096;; take = 140000) {        // (FEXTRA holds up to 65535 octets; names and comments are short)
      head.resize((size_t)std::min<uint64_t>(take, m.in_len));
      CK(ctx, cudaMemcpyAsync(head.data(), m.in, head.size(), cudaMemcpyDeviceToHost, st));
      CK(ctx, cudaStreamSynchronize(st));
      tbz_gzip_header h;
      if (tbz_gzip_header_parse(head.data(), head.size(), &h) != TBZ_OK) return TBZ_OK;
      if (h.verdict == TBZ_FINISHED) { hdr_bytes = h.header_len; break; }
      if (h.verdict != TBZ_INPUT_UNDERRUN || take != 4096 || head.size() == m.in_len) return TBZ_OK;
    }
  }
  const uint64_t trailer = fmt == TBZ_ZLIB ? 4 : fmt == TBZ_GZIP ? 8 : 0;
  if (m.in_len < hdr_bytes + trailer + 64 || m.in_len >= (1ull << 40) || m.out_cap >= (1ull << 32)) return TBZ_OK;
  const uint64_t mis = (uintptr_t)m.in & 3;
  const uint32_t *words = (const uint32_t *)(m.in - mis);
  const uint64_t end_bit = (mis + m.in_len) * 8;
  const uint64_t body_bit = (mis + hdr_bytes) * 8;
  // Chunk size, measured on a 1 GiB gzip member (r2u, r2w2, r2x2: 160 / 192 / 224 / 256 / 288 / 320 / 384 KiB ->
  // 17.1 / 18.4 / 15.4 / 19.0 / 20.3 / 16.0 / 23.1 ms): the block-start search costs per chunk and the decode's latency
  // grows with the chunk (one warp per chunk, one wave).  The scatter is one chunk: where the search's first hit in a
  // chunk is a false positive (about one chunk in a thousand), the chunk in front of it decodes twice as far and
  // the kernel waits for it (TBZ_KTIME=1 prints the slowest chunks).
  // Smaller members get smaller chunks — about a thousand of them, not below 64 KiB (256 MiB member, r2c256: 64 / 96 / 128 /
  // 160 / 224 KiB -> 5.85 / 5.49 / 5.63 / 5.72 / 6.50 ms: with few chunks the decode's latency is all there is)
  uint64_t chunk_bytes = kSplitChunkBytes;
  if (!getenv("TBZ_SPLIT_CHUNK_KB"))
    chunk_bytes = std::min<uint64_t>(kSplitChunkBytes, std::max<uint64_t>(64ull << 10, ((m.in_len - hdr_bytes) / 1000 + 63) & ~63ull));
  const uint64_t chunk_bits = chunk_bytes * 8;
  const uint32_t nchunks = (uint32_t)((end_bit - body_bit + chunk_bits - 1) / chunk_bits);
  if (nchunks < 4) return TBZ_OK;

  void *d_found = nullptr, *d_chunks = nullptr, *d_todo = nullptr, *d_cnt = nullptr, *d_slabs = nullptr, *d_sym = nullptr, *d_parts = nullptr;
  int32_t rc = TBZ_OK;
  auto cleanup = [&]() {
    dev_release(ctx, d_found); dev_release(ctx, d_chunks); dev_release(ctx, d_todo); dev_release(ctx, d_cnt);
    dev_release(ctx, d_slabs); dev_release(ctx, d_sym); dev_release(ctx, d_parts);
  };
#define SCK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { cleanup(); return fail(ctx, TBZ_E_CUDA, #call, e_); } } while (0)
#define SRC(x) do { rc = (x); if (rc) { cleanup(); return rc; } } while (0)
  SRC(dev_alloc(ctx, (size_t)(nchunks + 2) * 8, &d_found));
  SRC(dev_alloc(ctx, (size_t)nchunks * sizeof(Chunk), &d_chunks));
  SRC(dev_alloc(ctx, (size_t)nchunks * 4, &d_todo));
  SRC(dev_alloc(ctx, 256, &d_cnt));
  const size_t dec_smem = sizeof(tbzfast::WSmem) * tbzfast::WPC;
  // ---- K0: block starts
  SCK(cudaFuncSetAttribute(tbzsplit::k_split_find, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dec_smem));
  SCK(cudaMemsetAsync(d_found, 0xff, (size_t)nchunks * 8, st));
  tbzsplit::k_split_find<<<((uint64_t)nchunks * tbzsplit::FSUB + tbzfast::WPC - 1) / tbzfast::WPC, tbzfast::NT, dec_smem, st>>>(
      words, end_bit, body_bit, chunk_bits, nchunks, (uint64_t *)d_found);
  ctx->launches++;
  std::vector<uint64_t> found(nchunks);
  SCK(cudaMemcpyAsync(found.data(), d_found, (size_t)nchunks * 8, cudaMemcpyDeviceToHost, st));
  SCK(cudaStreamSynchronize(st));
  found[0] = body_bit;
  stage("find");
  // ---- K1: decode the chunks, validate the chain, repair it where a start was a false positive
  const uint64_t round_bytes = (uint64_t)tbzfast::NL * tbzfast::S_MAX / 8;
  const uint64_t slab_bytes = (uint64_t)tbzfast::SLAB_WORDS * 4;
  uint64_t want = 3 * (m.in_len / round_bytes) + 8ull * nchunks + 64;
  if (want * slab_bytes > (32ull << 30)) want = (32ull << 30) / slab_bytes;
  const uint32_t nslabs = (uint32_t)want;
  SRC(dev_alloc(ctx, (size_t)nslabs * slab_bytes, &d_slabs));
  SCK(cudaFuncSetAttribute(tbzsplit::k_split_decode, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dec_smem));
  std::vector<uint32_t> valid;                       // chunk indices that start a decode, ascending
  for (uint32_t c = 0; c < nchunks; c++) if (found[c] != tbzsplit::NONE64) valid.push_back(c);
  std::vector<Chunk> ch(nchunks);
  {
    // one pass: every candidate chunk decodes until it lands exactly on a later candidate start
    std::vector<unsigned long long> cands(valid.size());
    for (size_t v = 0; v < valid.size(); v++) {
      Chunk &c = ch[valid[v]];
      cands[v] = found[valid[v]];
      c.start_bit = found[valid[v]];
      c.stop_bit = v + 1 < valid.size() ? found[valid[v + 1]] : tbzsplit::NONE64;
      c.pad = (uint32_t)v;
    }
    void *d_cands = nullptr;
    SRC(dev_alloc(ctx, cands.size() * 8 + 8, &d_cands));
    SCK(cudaMemcpyAsync(d_cands, cands.data(), cands.size() * 8, cudaMemcpyHostToDevice, st));
    SCK(cudaMemcpyAsync(d_chunks, ch.data(), (size_t)nchunks * sizeof(Chunk), cudaMemcpyHostToDevice, st));
    SCK(cudaMemcpyAsync(d_todo, valid.data(), valid.size() * 4, cudaMemcpyHostToDevice, st));
    SCK(cudaMemsetAsync(d_cnt, 0, 256, st));
    int occ = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, tbzsplit::k_split_decode, tbzfast::NT, dec_smem);
    const int grid = (int)std::min<uint64_t>((valid.size() + tbzfast::WPC - 1) / tbzfast::WPC, (uint64_t)ctx->sm_count * std::max(1, occ));
    tbzsplit::k_split_decode<<<grid, tbzfast::NT, dec_smem, st>>>(words, end_bit, (Chunk *)d_chunks, (const uint32_t *)d_todo,
                                                                  (uint32_t)valid.size(), (uint32_t *)d_slabs, nslabs, (uint32_t *)d_cnt,
                                                                  (const unsigned long long *)d_cands, (uint32_t)cands.size());
    ctx->launches++;
    cudaError_t e1 = cudaMemcpyAsync(ch.data(), d_chunks, (size_t)nchunks * sizeof(Chunk), cudaMemcpyDeviceToHost, st);
    cudaError_t e2 = cudaStreamSynchronize(st);
    dev_release(ctx, d_cands);
    if (e1 != cudaSuccess || e2 != cudaSuccess) { cleanup(); return fail(ctx, TBZ_E_CUDA, "split decode", e1 != cudaSuccess ? e1 : e2); }
    // follow the landings from chunk 0; a chunk nobody lands on did not start at a block start
    std::vector<uint32_t> chain;
    size_t v = 0;
    for (;;) {
      const Chunk &c = ch[valid[v]];
      if (c.rec.status == 0) { cleanup(); return TBZ_OK; }      // a chunk on the chain did not decode: the sequential path owns the verdict
      chain.push_back(valid[v]);
      if (c.rec.status == 1) break;                               // the final block
      auto it = std::lower_bound(cands.begin() + v + 1, cands.end(), (unsigned long long)c.land_bit);
      if (it == cands.end() || *it != c.land_bit) { cleanup(); return TBZ_OK; }
      v = (size_t)(it - cands.begin());
    }
    if (ctx->ktime && !ctx->ktime_quiet) {            // which chunks held the kernel up
      std::vector<uint32_t> by(chain.begin(), chain.end());
      std::sort(by.begin(), by.end(), [&](uint32_t a, uint32_t b) { return ch[a].pad > ch[b].pad; });
      for (size_t k = 0; k < std::min<size_t>(4, by.size()); k++) {
        const Chunk &c = ch[by[k]];
        fprintf(stderr, "[tbz split]   slowest decode: chunk %u, %.0f kcycles, %.0f KiB in, %u bytes out\n", by[k], c.pad * 1.024,
                (double)(c.land_bit - c.start_bit) / 8192.0, c.rec.out_len);
      }
      const Chunk &c = ch[by[by.size() / 2]];
      fprintf(stderr, "[tbz split]   median decode: chunk %u, %.0f kcycles, %.0f KiB in, %u bytes out\n", by[by.size() / 2], c.pad * 1.024,
              (double)(c.land_bit - c.start_bit) / 8192.0, c.rec.out_len);
      // the candidates nobody landed on (false positives of the search): what their headers claim
      size_t ci = 0, shown = 0;
      for (uint32_t vc : valid) {
        if (ci < chain.size() && chain[ci] == vc) { ci++; continue; }
        if (shown++ >= 8) break;
        const uint64_t bit = ch[vc].start_bit;
        uint8_t raw[16] = {0};
        cudaMemcpy(raw, (const uint8_t *)words + (bit >> 3), 12, cudaMemcpyDeviceToHost);
        unsigned long long w = 0;
        for (int k = 7; k >= 0; k--) w = (w << 8) | raw[k];
        w >>= (bit & 7);
        fprintf(stderr, "[tbz split]   dropped candidate: chunk %u at bit %llu (+%llu in its chunk): final %llu hlit %llu hdist %llu hclen %llu, decode status %u\n",
                vc, (unsigned long long)bit, (unsigned long long)(bit - (body_bit + chunk_bits * vc)), w & 1, ((w >> 3) & 31) + 257,
                ((w >> 8) & 31) + 1, ((w >> 13) & 15) + 4, ch[vc].rec.status);
      }
    }
    valid.swap(chain);
  }
  stage("decode");
  // ---- output offsets
  std::vector<Chunk> vch(valid.size());
  uint64_t total = 0;
  for (size_t v = 0; v < valid.size(); v++) { vch[v] = ch[valid[v]]; vch[v].out_off = total; vch[v].ok = 0; total += vch[v].rec.out_len; }
  if (total > m.out_cap || total >= (1ull << 32)) { cleanup(); return TBZ_OK; }
  const uint64_t end_pos_bit = vch.back().land_bit;
  const uint64_t trailer_byte = (end_pos_bit + 7) / 8;                 // relative to the aligned base
  if (trailer_byte + trailer > mis + m.in_len) { cleanup(); return TBZ_OK; }   // truncated trailer: sequential kernel
  const uint32_t nv = (uint32_t)vch.size();
  SCK(cudaMemcpyAsync(d_chunks, vch.data(), (size_t)nv * sizeof(Chunk), cudaMemcpyHostToDevice, st));
  // ---- K2: symbols
  SRC(dev_alloc(ctx, (size_t)total * 2 + 64, &d_sym));
  SCK(cudaFuncSetAttribute(tbzsplit::k_split_resolve, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(tbzsplit::SymSmem)));
  {
    int occ = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, tbzsplit::k_split_resolve, tbzres::NT, sizeof(tbzsplit::SymSmem));
    const int grid = (int)std::min<uint64_t>(nv, (uint64_t)ctx->sm_count * std::max(1, occ));
    SCK(cudaMemsetAsync(d_cnt, 0, 256, st));
    tbzsplit::k_split_resolve<<<grid, tbzres::NT, sizeof(tbzsplit::SymSmem), st>>>((Chunk *)d_chunks, nv, (const uint32_t *)d_slabs,
                                                                                   (uint16_t *)d_sym, (uint32_t *)d_cnt);
    ctx->launches++;
  }
  SCK(cudaMemcpyAsync(vch.data(), d_chunks, (size_t)nv * sizeof(Chunk), cudaMemcpyDeviceToHost, st));
  SCK(cudaStreamSynchronize(st));
  for (const Chunk &c : vch) if (!c.ok) { cleanup(); return TBZ_OK; }
  stage("resolve");
  // ---- K3, K4: bytes
  {
    static const bool seq_tails = getenv("TBZ_SEQ_TAILS") != nullptr;   // the one-CTA walk, for comparison
    void *d_maps = nullptr;
    const size_t map_bytes = (size_t)nv * tbzsplit::TAILW * 2;
    if (!seq_tails && nv > 1 && dev_alloc(ctx, 2 * map_bytes, &d_maps) == TBZ_OK) {
      uint16_t *ma = (uint16_t *)d_maps, *mb = ma + (size_t)nv * tbzsplit::TAILW;
      const dim3 grid(nv, tbzsplit::TAILW / 1024);
      tbzsplit::k_tail_init<<<grid, 256, 0, st>>>((const Chunk *)d_chunks, (const uint16_t *)d_sym, ma);
      ctx->launches++;
      for (uint32_t stride = 1; stride < nv; stride <<= 1) {
        tbzsplit::k_tail_compose<<<dim3(nv, tbzsplit::TAILW / 2048), 256, 0, st>>>(ma, mb, stride);
        ctx->launches++;
        std::swap(ma, mb);
      }
      tbzsplit::k_tail_write<<<grid, 256, 0, st>>>((const Chunk *)d_chunks, ma, m.out);
      cudaError_t e_ = cudaStreamSynchronize(st);
      dev_release(ctx, d_maps);
      if (e_ != cudaSuccess) { cleanup(); return fail(ctx, TBZ_E_CUDA, "tail scan", e_); }
    } else {
      SCK(cudaFuncSetAttribute(tbzsplit::k_split_tails, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(tbzsplit::TailSmem)));
      tbzsplit::k_split_tails<<<1, 1024, sizeof(tbzsplit::TailSmem), st>>>((const Chunk *)d_chunks, nv, (const uint16_t *)d_sym, m.out);
    }
  }
  stage("tails");
  {
    std::vector<uint64_t> offs(nv + 1);
    for (uint32_t v = 0; v < nv; v++) offs[v] = vch[v].out_off;
    offs[nv] = total;
    SCK(cudaMemcpyAsync(d_found, offs.data(), (size_t)(nv + 1) * 8, cudaMemcpyHostToDevice, st));   // nv + 1 <= nchunks + 1 words: fits
    SCK(cudaStreamSynchronize(st));
    tbzsplit::k_split_translate<<<ctx->sm_count * 16, 256, 0, st>>>((const uint64_t *)d_found, nv, (const uint16_t *)d_sym, m.out, total);
  }
  ctx->launches += 2;
  stage("translate");
  // ---- K5: checksum, trailer (zlib.lisp:80-96, gzip.lisp:82-106)
  uint32_t ck = 0;
  uint8_t tr[8] = {0};
  if (trailer) {
    // gzip: the CTA-per-span CRC of the batched path (inflate_crc.cuh) over 1 MiB pieces; zlib: sums per 64 KiB segment
    const uint64_t seg_bytes = fmt == TBZ_GZIP ? (1ull << 20) : (uint64_t)tbzsplit::ASEG;
    const uint64_t nseg = (total + seg_bytes - 1) / seg_bytes;
    SRC(dev_alloc(ctx, (size_t)std::max<uint64_t>(1, nseg) * 8, &d_parts));
    if (nseg) {
      if (fmt == TBZ_GZIP) {
        SCK(cudaFuncSetAttribute(tbzcrc::k_span_crc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tbzcrc::SMEM_BYTES));
        tbzcrc::k_span_crc<<<(uint32_t)std::min<uint64_t>(nseg, (uint64_t)ctx->sm_count), tbzcrc::NT, tbzcrc::SMEM_BYTES, st>>>(
            m.out, total, (uint32_t)seg_bytes, (uint32_t *)d_parts);
      } else {
        tbzsplit::k_split_adler<<<(uint32_t)((nseg + 7) / 8), 256, 0, st>>>(m.out, total, (uint32_t *)d_parts);
      }
      ctx->launches++;
    }
    std::vector<uint32_t> parts(2 * std::max<uint64_t>(1, nseg));
    SCK(cudaMemcpyAsync(parts.data(), d_parts, (size_t)nseg * (fmt == TBZ_GZIP ? 4 : 8), cudaMemcpyDeviceToHost, st));
    SCK(cudaMemcpyAsync(tr, (const uint8_t *)words + trailer_byte, trailer, cudaMemcpyDeviceToHost, st));
    SCK(cudaStreamSynchronize(st));
    if (fmt == TBZ_GZIP) {
      // crc(A || B) = crc(A) * x^(8 |B|) + crc(B)
      const uint32_t xs = crc_x8n_h(seg_bytes);
      uint32_t c = 0;
      for (uint64_t s2 = 0; s2 < nseg; s2++) {
        const uint64_t len = std::min<uint64_t>(seg_bytes, total - s2 * seg_bytes);
        c = crc_mulmod_h(len == seg_bytes ? xs : crc_x8n_h(len), c) ^ parts[s2];
      }
      ck = c;
    } else {
      uint64_t s1 = 1, s2v = 0;
      for (uint64_t s = 0; s < nseg; s++) {
        const uint64_t len = std::min<uint64_t>(seg_bytes, total - s * seg_bytes);
        s2v = (s2v + len * s1 + parts[2 * s + 1]) % TBZ_ADLER_MOD;
        s1 = (s1 + parts[2 * s]) % TBZ_ADLER_MOD;
      }
      ck = (uint32_t)(s1 | (s2v << 16));
    }
  } else {
    SCK(cudaStreamSynchronize(st));
  }
  tbz_result r{};
  r.out_len = total;
  r.checksum = ck;
  r.where = TBZ_AT_BODY;
  r.path = 2;
  r.verdict = TBZ_FINISHED;
  r.in_used = trailer_byte + trailer - mis;
  if (fmt == TBZ_ZLIB) {
    const uint32_t t = ((uint32_t)tr[0] << 24) | ((uint32_t)tr[1] << 16) | ((uint32_t)tr[2] << 8) | tr[3];
    if (t != ck) r.verdict = TBZ_ERR_CHECKSUM;
  } else if (fmt == TBZ_GZIP) {
    const uint32_t t = tr[0] | ((uint32_t)tr[1] << 8) | ((uint32_t)tr[2] << 16) | ((uint32_t)tr[3] << 24);
    if (t != ck) { r.verdict = TBZ_ERR_CHECKSUM; r.in_used -= 4; }
  }
  SCK(cudaMemcpyAsync(d_result, &r, sizeof r, cudaMemcpyHostToDevice, st));
  SCK(cudaStreamSynchronize(st));
  stage("checksum");
  if (ctx->ktime && !ctx->ktime_quiet) fprintf(stderr, "[tbz split] %u chunks of %u, %llu bytes out\n", nv, nchunks, (unsigned long long)total);
  cleanup();
#undef SCK
#undef SRC
  *handled = true;
  return TBZ_OK;
}

// a sub-batch's result records, device -> mapped pinned host memory (see tbz_batch_launch)
static __global__ void k_results_to_host(const uint32_t *__restrict__ src, uint32_t *__restrict__ dst, uint32_t words) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < words) dst[i] = src[i];
}

static int32_t launch_kernels(tbz_batch *b) {
  tbz_ctx *ctx = b->ctx;
  if (!b->n) return TBZ_OK;
  uint32_t n = (uint32_t)b->n;
  if (b->fast_grid) {
    CK(ctx, cudaMemsetAsync(b->d_counters, 0, 256, ctx->stream));
    if (ctx->ktime) CK(ctx, cudaEventRecord(ctx->kev[0], ctx->stream));
    const size_t dec_smem = sizeof(tbzhd::WSmem) * tbzhd::WPC;
    CK(ctx, cudaFuncSetAttribute(k_inflate_decode, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dec_smem));
    k_inflate_decode<<<b->fast_grid, tbzhd::NT, dec_smem, ctx->stream>>>(
        (const DMember *)b->d_members, n, b->format, (tbzfast::P1Rec *)b->d_recs,
        (uint32_t *)b->d_slabs, b->nslabs, (uint32_t *)b->d_counters, (uint32_t *)b->d_todo, (unsigned char *)b->d_scratch);
    ctx->launches++;
    CK(ctx, cudaGetLastError());
    if (ctx->ktime) CK(ctx, cudaEventRecord(ctx->kev[1], ctx->stream));
    CK(ctx, cudaFuncSetAttribute(k_inflate_resolve, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(tbzp2::Smem)));
    k_inflate_resolve<<<b->res_grid, tbzp2::NT, sizeof(tbzp2::Smem), ctx->stream>>>(
        (const DMember *)b->d_members, (tbz_result *)b->d_results, n, b->format,
        (const tbzfast::P1Rec *)b->d_recs, (const uint32_t *)b->d_slabs, (uint32_t *)b->d_counters, (uint32_t *)b->d_todo);
    ctx->launches++;
    CK(ctx, cudaGetLastError());
    if (b->format == TBZ_GZIP) {   // gzip: CRC-32 of the finished members + trailer compare
      const int crc_grid = (int)std::min<uint64_t>(n, (uint64_t)ctx->sm_count);
      CK(ctx, cudaFuncSetAttribute(tbzcrc::k_member_crc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tbzcrc::SMEM_BYTES));
      tbzcrc::k_member_crc<<<crc_grid, tbzcrc::NT, tbzcrc::SMEM_BYTES, ctx->stream>>>(
          (const DMember *)b->d_members, (tbz_result *)b->d_results, n, (const tbzfast::P1Rec *)b->d_recs,
          (uint32_t *)b->d_counters, (uint32_t *)b->d_todo);
      ctx->launches++;
      CK(ctx, cudaGetLastError());
    }
    if (ctx->ktime) CK(ctx, cudaEventRecord(ctx->kev[2], ctx->stream));
    k_inflate_seq<<<(n + SEQ_WARPS - 1) / SEQ_WARPS, SEQ_WARPS * 32, 0, ctx->stream>>>(
        (const DMember *)b->d_members, (tbz_result *)b->d_results, n, b->format,
        (const uint32_t *)b->d_todo, (const uint32_t *)b->d_counters + 1);
    ctx->launches++;
    CK(ctx, cudaGetLastError());
    if (ctx->ktime) {
      CK(ctx, cudaEventRecord(ctx->kev[3], ctx->stream));
      CK(ctx, cudaEventSynchronize(ctx->kev[3]));
      float a = 0, c = 0, d = 0;
      cudaEventElapsedTime(&a, ctx->kev[0], ctx->kev[1]);
      cudaEventElapsedTime(&c, ctx->kev[1], ctx->kev[2]);
      cudaEventElapsedTime(&d, ctx->kev[2], ctx->kev[3]);
      uint32_t cnt[4] = {0, 0, 0, 0};
      cudaMemcpy(cnt, b->d_counters, sizeof cnt, cudaMemcpyDeviceToHost);
      ctx->last_kms[0] = a; ctx->last_kms[1] = c; ctx->last_kms[2] = d;
      if (!ctx->ktime_quiet) fprintf(stderr, "[tbz] decode %.3f ms, resolve %.3f ms, seq %.3f ms (%u members), %u slabs\n", a, c, d, cnt[1], cnt[2]);
    }
    return TBZ_OK;
  }
  k_inflate_seq<<<(n + SEQ_WARPS - 1) / SEQ_WARPS, SEQ_WARPS * 32, 0, ctx->stream>>>(
      (const DMember *)b->d_members, (tbz_result *)b->d_results, n, b->format, nullptr, nullptr);
  ctx->launches++;
  CK(ctx, cudaGetLastError());
  return TBZ_OK;
}

extern "C" int32_t tbz_batch_launch(tbz_batch *b) {
  if (!b) return TBZ_E_ARG;
  tbz_ctx *ctx = b->ctx;
  CK(ctx, cudaSetDevice(ctx->device));
  struct StreamScope {                                 // everything below enqueues on the batch's stream
    tbz_ctx *c; cudaStream_t saved;
    StreamScope(tbz_ctx *c_, cudaStream_t s) : c(c_), saved(c_->stream) { c->stream = s; }
    ~StreamScope() { c->stream = saved; }
  } scope(ctx, b->stream);
  if (!b->device_ptrs && b->n) {
    if (b->in_direct) {
      CK(ctx, cudaMemcpyAsync(b->d_in, b->in_span, b->in_total, cudaMemcpyHostToDevice, ctx->stream));
    } else {
      int32_t rc = ensure_stage(ctx, &ctx->stage_in, &ctx->stage_in_cap, b->in_total);
      if (rc) return rc;
      CK(ctx, cudaStreamSynchronize(ctx->stream));   // staging buffer may still be in flight
      uint8_t *st = (uint8_t *)ctx->stage_in;
      parallel_for(b->n, b->in_total, [&](size_t lo, size_t hi) {
        for (size_t i = lo; i < hi; i++) memcpy(st + b->in_off[i], b->host[i].in, b->host[i].in_len);
      });
      CK(ctx, cudaMemcpyAsync(b->d_in, st, b->in_total, cudaMemcpyHostToDevice, ctx->stream));
    }
  }
  CK(ctx, cudaEventRecord(b->ev0, ctx->stream));
  int32_t rc = launch_kernels(b);
  if (rc) return rc;
  for (auto &bm : b->big) {                            // large members: split across the GPU, else the sequential kernel
    bool handled = false;
    rc = split_inflate(ctx, b->format, bm.second, (tbz_result *)b->d_results + bm.first, &handled);
    if (rc) return rc;
    if (!handled) {
      CK(ctx, cudaMemcpyAsync((DMember *)b->d_members + bm.first, &bm.second, sizeof(DMember), cudaMemcpyHostToDevice, ctx->stream));
      const uint32_t one = (uint32_t)bm.first;
      void *d_one = nullptr;
      rc = dev_alloc(ctx, 256, &d_one);
      if (rc) return rc;
      const uint32_t hdr[2] = {one, 1u};
      CK(ctx, cudaMemcpyAsync(d_one, hdr, sizeof hdr, cudaMemcpyHostToDevice, ctx->stream));
      k_inflate_seq<<<1, SEQ_WARPS * 32, 0, ctx->stream>>>((const DMember *)b->d_members, (tbz_result *)b->d_results, 1, b->format,
                                                           (const uint32_t *)d_one, (const uint32_t *)d_one + 1);
      ctx->launches++;
      CK(ctx, cudaStreamSynchronize(ctx->stream));
      DMember z = bm.second; z.in_len = 0; z.out_cap = 0;
      CK(ctx, cudaMemcpyAsync((DMember *)b->d_members + bm.first, &z, sizeof(DMember), cudaMemcpyHostToDevice, ctx->stream));
      CK(ctx, cudaStreamSynchronize(ctx->stream));
      dev_release(ctx, d_one);
    }
  }
  CK(ctx, cudaEventRecord(b->ev1, ctx->stream));
  if (b->eager_res && b->n) {
    // The result records go to the host's pinned memory by a kernel's stores, not through a copy engine: as a DMA they
    // queued behind the 30 MB output copies of the parts in front (0.5 ms each), and the kernels of the next part on
    // this stream waited with them (TBZ_PIPE_TRACE=1 shows the timeline)
    static const bool by_dma = getenv("TBZ_PIPE_RESULTS_DMA") != nullptr;
    void *host_dev = nullptr;
    if (!by_dma && cudaHostGetDevicePointer(&host_dev, b->eager_res, 0) == cudaSuccess && host_dev) {
      const uint32_t words = (uint32_t)(b->n * sizeof(tbz_result) / 4);
      k_results_to_host<<<(words + 255) / 256, 256, 0, ctx->stream>>>((const uint32_t *)b->d_results, (uint32_t *)host_dev, words);
      ctx->launches++;
    } else {
      cudaGetLastError();
      CK(ctx, cudaMemcpyAsync(b->eager_res, b->d_results, b->n * sizeof(tbz_result), cudaMemcpyDeviceToHost, ctx->stream));
    }
    CK(ctx, cudaEventRecord(b->ev_res, ctx->stream));
  }
  b->launched = true;
  return TBZ_OK;
}

// Outputs that are exactly adjacent in the caller's memory leave the device straight into it: of every member the
// bytes it PRODUCED, never its capacity (the reference does not touch an output buffer past the count it returns,
// api.lisp:35-61, and what lies there on the device is whatever an earlier batch left).  Members that filled their
// buffers merge with their right neighbour into one copy: a batch of exact-size outputs is one DMA.
static int32_t copy_out_direct(tbz_batch *b, const tbz_result *res, cudaStream_t st) {
  tbz_ctx *ctx = b->ctx;
  uint64_t i = 0;
  while (i < b->n) {
    uint64_t j = i, bytes = 0;
    for (;;) {
      const uint64_t got = std::min<uint64_t>(res[j].out_len, b->host[j].out_cap);
      bytes += got;
      if (got != b->host[j].out_cap || j + 1 == b->n) break;
      j++;
    }
    if (bytes)
      CK(ctx, cudaMemcpyAsync(b->host[i].out, (const uint8_t *)b->d_out + b->out_off[i], bytes, cudaMemcpyDeviceToHost, st));
    i = j + 1;
  }
  return TBZ_OK;
}

extern "C" int32_t tbz_batch_finish(tbz_batch *b, tbz_result *r) {
  if (!b) return TBZ_E_ARG;
  tbz_ctx *ctx = b->ctx;
  CK(ctx, cudaSetDevice(ctx->device));
  if (!b->launched) return fail(ctx, TBZ_E_STATE, "tbz_batch_finish before tbz_batch_launch");
  if (!b->n) return TBZ_OK;
  cudaStream_t st = b->stream;
  if (b->eager_res) {                                   // (the pipelined path: results are on their way to pinned memory)
    CK(ctx, cudaStreamSynchronize(st));
    if (!b->copies_issued) { int32_t rc = copy_out_direct(b, b->eager_res, st); if (rc) return rc; }
    CK(ctx, cudaStreamSynchronize(st));
    return TBZ_OK;
  }
  std::vector<tbz_result> tmp;
  tbz_result *res = r;
  if (!res) { tmp.resize(b->n); res = tmp.data(); }
  CK(ctx, cudaMemcpyAsync(res, b->d_results, b->n * sizeof(tbz_result), cudaMemcpyDeviceToHost, st));
  if (!b->device_ptrs) {
    if (b->out_direct) {
      CK(ctx, cudaStreamSynchronize(st));               // the results say how much every member produced
      int32_t rc = copy_out_direct(b, res, st);
      if (rc) return rc;
      CK(ctx, cudaStreamSynchronize(st));
    } else {
      int32_t rc = ensure_stage(ctx, &ctx->stage_out, &ctx->stage_out_cap, b->out_total);
      if (rc) return rc;
      CK(ctx, cudaMemcpyAsync(ctx->stage_out, b->d_out, b->out_total, cudaMemcpyDeviceToHost, st));
      CK(ctx, cudaStreamSynchronize(st));
      const uint8_t *st = (const uint8_t *)ctx->stage_out;
      parallel_for(b->n, b->out_total, [&](size_t lo, size_t hi) {
        for (size_t i = lo; i < hi; i++)
          memcpy(b->host[i].out, st + b->out_off[i], std::min<uint64_t>(res[i].out_len, b->host[i].out_cap));
      });
    }
  } else {
    CK(ctx, cudaStreamSynchronize(st));
  }
  return TBZ_OK;
}

static int32_t batch_device_ms(tbz_batch *b, float *ms) {
  tbz_ctx *ctx = b->ctx;
  if (!ms) return TBZ_OK;
  *ms = 0.f;
  if (!b->n) return TBZ_OK;
  CK(ctx, cudaEventSynchronize(b->ev1));
  CK(ctx, cudaEventElapsedTime(ms, b->ev0, b->ev1));
  return TBZ_OK;
}

// Host buffers, many members: the batch is cut into contiguous sub-batches that travel down several
// streams, so the H2D copy of one, the kernels of another and the D2H copy of a third overlap (the
// copy engines and the SMs are separate units).  Needs dense inputs and adjacent outputs (the DMA
// then runs straight on the caller's memory); anything else takes the one-piece path.
static int32_t inflate_batch_pipelined(tbz_ctx *ctx, int32_t format, const tbz_member *m, uint64_t n,
                                       tbz_result *r, uint32_t flags, float *device_ms, bool *done) {
  *done = false;
  uint64_t bytes = 0;
  for (uint64_t i = 0; i < n; i++) {
    if (m[i].in_len >= kSplitMinBytes) return TBZ_OK;
    bytes += m[i].in_len + m[i].out_cap;
  }
  if (bytes < (32ull << 20)) return TBZ_OK;
  // (measured, config 2 on one B200, r2aa: parts x streams 4x3 42.9 GB/s, 8x3 43.4, 12x3 41.1, 12x6 37.4, 24x6 37.1
  //  against 54.1 for the same bytes as plain copies: few, large DMAs)
  static const uint64_t max_parts = getenv("TBZ_PIPE_PARTS") ? strtoull(getenv("TBZ_PIPE_PARTS"), nullptr, 10) : 8;
  static const uint64_t npipe = std::min<uint64_t>(tbz_ctx::kPipeStreams - 2, getenv("TBZ_PIPE_STREAMS") ? std::max<uint64_t>(1, strtoull(getenv("TBZ_PIPE_STREAMS"), nullptr, 10)) : 3);
  // The first output byte can leave only after one part has been uploaded, decoded (a member's decode latency is the
  // same ~0.5 ms whatever the part's size) and resolved, and from then on the D2H engine is the bottleneck: the first
  // parts are small (a sixteenth of the batch each), so that the wait in front of the first DMA holds little work
  static const uint64_t first_div = getenv("TBZ_PIPE_FIRST") ? strtoull(getenv("TBZ_PIPE_FIRST"), nullptr, 10) : 16;
  std::vector<uint64_t> cut;                            // part p = members [cut[p], cut[p + 1])
  {
    const uint64_t eq = std::min<uint64_t>(max_parts, std::max<uint64_t>(2, n / 256));
    cut.push_back(0);
    uint64_t at = 0;
    if (first_div > eq && n / first_div >= 128) { at = n / first_div; cut.push_back(at); at *= 2; cut.push_back(at); }
    for (uint64_t p = 1; p <= eq; p++) cut.push_back(at + (n - at) * p / eq);
  }
  const uint64_t parts = cut.size() - 1;
  std::vector<tbz_batch *> sub(parts, nullptr);
  {
    int32_t src = ensure_stage(ctx, &ctx->stage_res, &ctx->stage_res_cap, n * sizeof(tbz_result));
    if (src) return src;
  }
  tbz_result *pinned = (tbz_result *)ctx->stage_res;
  cudaStream_t saved = ctx->stream;
  int32_t rc = TBZ_OK;
  bool ok = true;
  static const bool trace = getenv("TBZ_PIPE_TRACE") != nullptr;   // the host's timeline of the pipeline (TBZ_KTIME would add its own syncs)
  const auto t_begin = std::chrono::steady_clock::now();
  auto since = [&]() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_begin).count(); };
  for (uint64_t p = 0; p < parts && !rc; p++) {
    const uint64_t lo = cut[p], hi = cut[p + 1];
    ctx->stream = ctx->pstream[p % npipe];
    rc = tbz_batch_prepare(ctx, format, m + lo, hi - lo, flags, &sub[p]);
    if (!rc && !(sub[p]->in_direct && sub[p]->out_direct)) { ok = false; break; }
    if (!rc) { sub[p]->eager_res = pinned + lo; rc = tbz_batch_launch(sub[p]); }
  }
  ctx->stream = saved;
  if (trace) fprintf(stderr, "[tbz pipe] %llu parts prepared and launched at %.3f ms\n", (unsigned long long)parts, since());
  float total_ms = 0.f;
  // part after part: its results arrive, the copies of what it produced go into its stream (they overlap the kernels
  // of the parts behind it) ...
  // (the copies get streams of their own: a part's stream already holds the kernels of the part queued behind it)
  for (uint64_t p = 0; p < parts && !rc && ok; p++) {
    if (!sub[p] || !sub[p]->launched) continue;
    cudaError_t e = cudaEventSynchronize(sub[p]->ev_res);
    if (e != cudaSuccess) { rc = fail(ctx, TBZ_E_CUDA, "pipelined results", e); break; }
    const double t_res = trace ? since() : 0.0;
    rc = copy_out_direct(sub[p], sub[p]->eager_res, ctx->pstream[tbz_ctx::kPipeStreams - 1 - (p & 1)]);
    sub[p]->copies_issued = true;
    if (trace) fprintf(stderr, "[tbz pipe]   part %llu (%llu members): results at %.3f ms, its D2H issued at %.3f ms\n", (unsigned long long)p,
                       (unsigned long long)(cut[p + 1] - cut[p]), t_res, since());
  }
  for (int k = 0; k < 2; k++) {
    cudaError_t e = cudaStreamSynchronize(ctx->pstream[tbz_ctx::kPipeStreams - 1 - k]);
    if (e != cudaSuccess && !rc) rc = fail(ctx, TBZ_E_CUDA, "pipelined D2H", e);
  }
  if (trace) fprintf(stderr, "[tbz pipe] all D2H done at %.3f ms\n", since());
  // ... then everything is waited for
  for (uint64_t p = 0; p < parts; p++) {
    if (!sub[p]) continue;
    if (!rc && ok && sub[p]->launched) {
      rc = tbz_batch_finish(sub[p], nullptr);
      float ms = 0.f;
      if (!rc) rc = batch_device_ms(sub[p], &ms);
      total_ms += ms;
    }
    tbz_batch_destroy(sub[p]);
  }
  if (rc) return rc;
  if (!ok) return TBZ_OK;                              // (sub-batches already launched rewrote nothing the one-piece path will not rewrite)
  memcpy(r, pinned, n * sizeof(tbz_result));
  if (device_ms) *device_ms = total_ms;
  if (trace) fprintf(stderr, "[tbz pipe] finished at %.3f ms\n", since());
  *done = true;
  return TBZ_OK;
}

extern "C" int32_t tbz_inflate_batch(tbz_ctx *ctx, int32_t format, const tbz_member *m, uint64_t n,
                                     tbz_result *r, uint32_t flags, float *device_ms) {
  if (ctx && m && r && n >= 512 && !(flags & TBZ_FLAG_DEVICE_PTRS)) {
    bool done = false;
    int32_t prc = inflate_batch_pipelined(ctx, format, m, n, r, flags, device_ms, &done);
    if (prc || done) return prc;
  }
  tbz_batch *b = nullptr;
  int32_t rc = tbz_batch_prepare(ctx, format, m, n, flags, &b);
  if (rc) return rc;
  rc = tbz_batch_launch(b);
  if (!rc) rc = tbz_batch_finish(b, r);
  if (!rc) rc = batch_device_ms(b, device_ms);
  tbz_batch_destroy(b);
  return rc;
}

extern "C" int32_t tbz_inflate_single(tbz_ctx *ctx, int32_t format, const uint8_t *in, uint64_t in_len,
                                      uint8_t *out, uint64_t out_cap, tbz_result *r, uint32_t flags,
                                      float *device_ms) {
  tbz_member m{in, in_len, out, out_cap};
  return tbz_inflate_batch(ctx, format, &m, 1, r, flags, device_ms);
}

// Decode a device-resident input fully into a device buffer that grows until the stream no longer overflows.
// tail4: the last four octets of the input (host copy; gzip's ISIZE is a sizing hint only: gzip.lisp:95-106 ignores it).
static int32_t inflate_resident(tbz_ctx *ctx, int32_t format, const void *d_in, uint64_t in_len, const uint8_t *tail4,
                                void **d_out, uint64_t *d_cap, tbz_result *res) {
  CK(ctx, cudaSetDevice(ctx->device));
  uint64_t cap = *d_cap;
  if (!*d_out) {
    cap = std::max<uint64_t>(65536, in_len * 4);
    if (format == TBZ_GZIP && in_len >= 18 && tail4) {
      uint32_t isz; memcpy(&isz, tail4, 4);
      if (isz >= cap / 8 && isz <= in_len * 1100 + 65536) cap = (uint64_t)isz + 64;
    }
  }
  int32_t rc;
  for (;;) {
    if (!*d_out) {
      rc = dev_alloc(ctx, cap, d_out);
      if (rc) return rc;
      *d_cap = cap;
    }
    tbz_member m{(const uint8_t *)d_in, in_len, (uint8_t *)*d_out, *d_cap};
    rc = tbz_inflate_batch(ctx, format, &m, 1, res, TBZ_FLAG_DEVICE_PTRS, nullptr);
    if (rc) break;
    if (res->verdict != TBZ_OUTPUT_OVERFLOW) break;
    dev_release(ctx, *d_out); *d_out = nullptr;
    cap = *d_cap * 4;
  }
  return rc;
}

// Decode `in` (host) fully into a device buffer that grows until the stream no longer overflows.
static int32_t inflate_to_device(tbz_ctx *ctx, int32_t format, const uint8_t *in, uint64_t in_len,
                                 void **d_out, uint64_t *d_cap, tbz_result *res) {
  CK(ctx, cudaSetDevice(ctx->device));
  void *d_in = nullptr;
  int32_t rc = dev_alloc(ctx, in_len + 16, &d_in);
  if (rc) return rc;
  if (in_len) {
    cudaError_t e = cudaMemcpyAsync(d_in, in, in_len, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) { dev_release(ctx, d_in); return fail(ctx, TBZ_E_CUDA, "H2D", e); }
  }
  rc = inflate_resident(ctx, format, d_in, in_len, in_len >= 4 ? in + in_len - 4 : nullptr, d_out, d_cap, res);
  dev_release(ctx, d_in);
  return rc;
}

extern "C" int32_t tbz_inflate_alloc(tbz_ctx *ctx, int32_t format, const uint8_t *in, uint64_t in_len,
                                     uint8_t **out, tbz_result *r) {
  if (!ctx || !out || !r || (in_len && !in)) return fail(ctx, TBZ_E_ARG, "tbz_inflate_alloc: bad argument");
  *out = nullptr;
  void *d_out = nullptr; uint64_t d_cap = 0;
  int32_t rc = inflate_to_device(ctx, format, in, in_len, &d_out, &d_cap, r);
  if (!rc) {
    uint8_t *h = (uint8_t *)malloc(r->out_len ? r->out_len : 1);
    if (!h) rc = fail(ctx, TBZ_E_NOMEM, "malloc");
    else {
      if (r->out_len) {
        cudaError_t e = cudaMemcpyAsync(h, d_out, r->out_len, cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) { free(h); h = nullptr; rc = fail(ctx, TBZ_E_CUDA, "D2H", e); }
      }
      *out = h;
    }
  }
  dev_release(ctx, d_out);
  return rc;
}

// =============================================================================================
// multi-GPU: host-side partition, one thread per device, no collective
// =============================================================================================
extern "C" int32_t tbz_partition(const uint64_t *in_len, uint64_t n, int32_t g, int32_t *owner) {
  if (g <= 0 || (n && (!in_len || !owner))) return TBZ_E_ARG;
  std::vector<uint64_t> idx(n);
  for (uint64_t i = 0; i < n; i++) idx[i] = i;
  std::stable_sort(idx.begin(), idx.end(), [&](uint64_t a, uint64_t b) { return in_len[a] > in_len[b]; });
  std::vector<uint64_t> load(g, 0), cnt(g, 0);
  for (uint64_t k = 0; k < n; k++) {
    int best = 0;
    for (int d = 1; d < g; d++)
      if (load[d] < load[best] || (load[d] == load[best] && cnt[d] < cnt[best])) best = d;
    owner[idx[k]] = best;
    load[best] += in_len[idx[k]] + 1;
    cnt[best]++;
  }
  return TBZ_OK;
}

extern "C" int32_t tbz_inflate_batch_multi(tbz_ctx *const *ctxs, int32_t g, int32_t format,
                                           const tbz_member *m, uint64_t n, tbz_result *r,
                                           uint32_t flags, float *device_ms_per_gpu) {
  if (!ctxs || g <= 0 || (n && (!m || !r))) return TBZ_E_ARG;
  if (flags & TBZ_FLAG_DEVICE_PTRS) return fail(ctxs[0], TBZ_E_ARG, "tbz_inflate_batch_multi takes host members");
  std::vector<uint64_t> lens(n);
  uint64_t total = 0, biggest = 0;
  for (uint64_t i = 0; i < n; i++) { lens[i] = m[i].in_len; total += lens[i] + 1; biggest = std::max(biggest, lens[i] + 1); }
  std::vector<int32_t> rcs(g, TBZ_OK);
  std::vector<std::thread> th;
  // Members of similar size (the common batch): consecutive ranges of equal compressed size.  A range keeps what the
  // caller's layout offers — dense inputs, adjacent outputs: the engine then DMAs straight from and to the caller's
  // memory, in pipelined parts — which a size-sorted assignment scatters.
  if (n >= (uint64_t)g && biggest * (uint64_t)g * 8 <= total) {
    std::vector<uint64_t> cut(g + 1, n);
    cut[0] = 0;
    uint64_t acc = 0; int d = 1;
    for (uint64_t i = 0; i < n && d < g; i++) {
      acc += lens[i] + 1;
      if (acc * g >= total * d) cut[d++] = i + 1;
    }
    for (int dd = 0; dd < g; dd++)
      th.emplace_back([&, dd] {
        float ms = 0.f;
        const uint64_t lo = cut[dd], hi = cut[dd + 1];
        if (hi > lo) rcs[dd] = tbz_inflate_batch(ctxs[dd], format, m + lo, hi - lo, r + lo, flags, &ms);
        if (device_ms_per_gpu) device_ms_per_gpu[dd] = ms;
      });
    for (auto &t : th) t.join();
    for (int dd = 0; dd < g; dd++) if (rcs[dd]) return rcs[dd];
    return TBZ_OK;
  }
  std::vector<int32_t> owner(n);
  int32_t rc = tbz_partition(lens.data(), n, g, owner.data());
  if (rc) return rc;
  std::vector<std::vector<uint64_t>> part(g);
  for (uint64_t i = 0; i < n; i++) part[owner[i]].push_back(i);
  for (int d = 0; d < g; d++)
    th.emplace_back([&, d] {
      std::vector<tbz_member> mm(part[d].size());
      std::vector<tbz_result> rr(part[d].size());
      for (size_t k = 0; k < mm.size(); k++) mm[k] = m[part[d][k]];
      float ms = 0.f;
      rcs[d] = tbz_inflate_batch(ctxs[d], format, mm.data(), mm.size(), rr.data(), flags, &ms);
      if (device_ms_per_gpu) device_ms_per_gpu[d] = ms;
      if (rcs[d] == TBZ_OK)
        for (size_t k = 0; k < mm.size(); k++) r[part[d][k]] = rr[k];
    });
  for (auto &t : th) t.join();
  for (int d = 0; d < g; d++) if (rcs[d]) return rcs[d];
  return TBZ_OK;
}

// =============================================================================================
// sessions: decompress / replace-output-buffer over a device-resident decoded member
// =============================================================================================
struct tbz_session {
  tbz_ctx *ctx;
  int format;
  void *d_in = nullptr; uint64_t in_cap = 0, in_len = 0;   // every octet handed over so far, on the device: a call uploads only its own
  uint8_t tail[4] = {0, 0, 0, 0}; // the last four of them (gzip ISIZE: a sizing hint)
  uint64_t last_n = 0;            // octets of the last call
  bool resumable = false;         // the caller feeds pieces: calls resume at block boundaries (k_inflate_session)
  void *d_rs = nullptr, *d_res = nullptr;   // tbzseq::Resume and the result record of the resumable kernel
  bool decoded = false;           // d_out/total reflect the input
  void *d_out = nullptr; uint64_t d_cap = 0;
  tbz_result total{};
  uint64_t served = 0;            // decoded bytes already delivered
  uint8_t *out = nullptr; uint64_t cap = 0, off = 0;
  bool finished = false, underrun = false, overflow = false;
  int32_t error = 0;
};

extern "C" int32_t tbz_session_create(tbz_ctx *ctx, int32_t format, tbz_session **s) {
  if (!ctx || !s || format < 0 || format > 2) return TBZ_E_ARG;
  tbz_session *x = new tbz_session();
  x->ctx = ctx; x->format = format;
  *s = x;
  return TBZ_OK;
}
extern "C" int32_t tbz_session_destroy(tbz_session *s) {
  if (!s) return TBZ_OK;
  dev_release(s->ctx, s->d_out);
  dev_release(s->ctx, s->d_in);
  dev_release(s->ctx, s->d_rs); dev_release(s->ctx, s->d_res);
  delete s;
  return TBZ_OK;
}
extern "C" int32_t tbz_session_set_output(tbz_session *s, uint8_t *out, uint64_t cap) {
  if (!s || (cap && !out)) return TBZ_E_ARG;
  s->out = out; s->cap = cap; s->off = 0;
  return TBZ_OK;
}
extern "C" int32_t tbz_session_rebind_output(tbz_session *s, uint8_t *out) {
  if (!s || (s->cap && !out)) return TBZ_E_ARG;
  s->out = out;
  return TBZ_OK;
}
extern "C" int32_t tbz_session_replace_output(tbz_session *s, uint8_t *out, uint64_t cap) {
  if (!s || (cap && !out)) return TBZ_E_ARG;
  if (!(s->off == 0 || s->overflow)) return TBZ_E_BUFFER_SWITCH;     // api.lisp:13-18
  s->out = out; s->cap = cap; s->off = 0; s->overflow = false;
  return TBZ_OK;
}
extern "C" int32_t tbz_session_flags(tbz_session *s, int32_t *fin, int32_t *under, int32_t *over) {
  if (!s) return TBZ_E_ARG;
  if (fin) *fin = s->finished;
  if (under) *under = s->underrun;
  if (over) *over = s->overflow;
  return TBZ_OK;
}

// A call of a session that is fed in pieces: one warp decodes from the last block boundary an earlier call reached
// (deflate.lisp:65-88,114-137 saves its whole state machine instead) over everything that has arrived since, into the
// session's device buffer (the window of the blocks before it is there), and leaves the new boundary and the running
// checksum behind.  The cost of a call is the block it is in plus the new octets, not the stream so far.
static int32_t session_resume(tbz_session *s) {
  tbz_ctx *ctx = s->ctx;
  CK(ctx, cudaSetDevice(ctx->device));
  int32_t rc;
  if (!s->d_rs) {
    rc = dev_alloc(ctx, 256, &s->d_rs);
    if (!rc) rc = dev_alloc(ctx, 256, &s->d_res);
    if (rc) return rc;
    CK(ctx, cudaMemsetAsync(s->d_rs, 0, 256, ctx->stream));
  }
  if (!s->d_out) {
    s->d_cap = std::max<uint64_t>(1 << 20, s->in_len * 4);
    rc = dev_alloc(ctx, s->d_cap, &s->d_out);
    if (rc) return rc;
  }
  for (;;) {
    DMember m{(const uint8_t *)s->d_in, s->in_len, (uint8_t *)s->d_out, s->d_cap};
    k_inflate_session<<<1, 32, 0, ctx->stream>>>(m, s->format, (tbz_result *)s->d_res, (tbzseq::Resume *)s->d_rs);
    ctx->launches++;
    CK(ctx, cudaGetLastError());
    CK(ctx, cudaMemcpyAsync(&s->total, s->d_res, sizeof(tbz_result), cudaMemcpyDeviceToHost, ctx->stream));
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    if (s->total.verdict != TBZ_OUTPUT_OVERFLOW) break;
    // the session's own buffer is full (not the caller's): a larger one, the bytes so far stay (they are the window)
    void *d = nullptr;
    const uint64_t cap = s->d_cap * 4;
    rc = dev_alloc(ctx, cap, &d);
    if (rc) return rc;
    CK(ctx, cudaMemcpyAsync(d, s->d_out, s->d_cap, cudaMemcpyDeviceToDevice, ctx->stream));
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    dev_release(ctx, s->d_out);
    s->d_out = d; s->d_cap = cap;
  }
  return TBZ_OK;
}

extern "C" int32_t tbz_session_consumed(tbz_session *s, uint64_t *n) {
  if (!s || !n) return TBZ_E_ARG;
  *n = s->last_n;
  if (s->finished && s->decoded) {
    const uint64_t before = s->in_len - s->last_n;                    // octets of the earlier calls
    *n = s->total.in_used > before ? std::min<uint64_t>(s->total.in_used - before, s->last_n) : 0;
  }
  return TBZ_OK;
}

extern "C" int32_t tbz_session_decompress(tbz_session *s, const uint8_t *in, uint64_t n,
                                          int64_t *ret, int32_t *verdict) {
  if (!s || !ret || !verdict || (n && !in)) return TBZ_E_ARG;
  tbz_ctx *ctx = s->ctx;
  if (s->error) { *ret = -1; *verdict = s->error; return TBZ_E_STATE; }
  if (s->finished) { *ret = -1; *verdict = TBZ_FINISHED; return TBZ_E_STATE; }   // ecase on :done (gzip.lisp:279-286)
  s->underrun = false;                       // deflate.lisp:102-103
  s->last_n = n;
  if (n) {
    // the new octets join the ones already on the device (the buffer doubles when it is full)
    CK(ctx, cudaSetDevice(ctx->device));
    if (s->in_len + n + 16 > s->in_cap) {
      const uint64_t cap = std::max<uint64_t>(std::max<uint64_t>(65536, 2 * s->in_cap), s->in_len + n + 16);
      void *d = nullptr;
      int32_t rc = dev_alloc(ctx, cap, &d);
      if (rc) return rc;
      if (s->in_len) CK(ctx, cudaMemcpyAsync(d, s->d_in, s->in_len, cudaMemcpyDeviceToDevice, ctx->stream));
      CK(ctx, cudaStreamSynchronize(ctx->stream));
      dev_release(ctx, s->d_in);
      s->d_in = d; s->in_cap = cap;
    }
    cudaError_t e = cudaMemcpyAsync((uint8_t *)s->d_in + s->in_len, in, n, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) return fail(ctx, TBZ_E_CUDA, "session H2D", e);
    for (uint64_t i = 0; i < 4; i++) s->tail[i] = i + n >= 4 ? in[n - 4 + i] : s->tail[i + n];   // (shift in the new octets)
    s->in_len += n;
    s->decoded = false;
  }
  if (!s->decoded) {
    if (!s->d_in) { int32_t rc = dev_alloc(ctx, 65536, &s->d_in); if (rc) return rc; s->in_cap = 65536; }
    int32_t rc;
    if (!s->resumable) {
      // the first look at the stream: the whole engine (fast kernels, split decode).  A stream that is complete —
      // the usual call: a context over all of it — is done here.
      rc = inflate_resident(ctx, s->format, s->d_in, s->in_len, s->in_len >= 4 ? s->tail : nullptr, &s->d_out, &s->d_cap, &s->total);
      if (rc) return rc;
      if (s->total.verdict == TBZ_INPUT_UNDERRUN) s->resumable = true;   // pieces: from the next call on, resume at block boundaries
    } else {
      rc = session_resume(s);
      if (rc) return rc;
    }
    s->decoded = true;
  }
  uint64_t room = s->cap - s->off, left = s->total.out_len - s->served;
  uint64_t take = std::min(room, left);
  if (take) {
    cudaError_t e = cudaMemcpyAsync(s->out + s->off, (const uint8_t *)s->d_out + s->served, take,
                                    cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) return fail(ctx, TBZ_E_CUDA, "session D2H", e);
    s->off += take; s->served += take;
  }
  if (left > room) {                          // every overflow site fires with the buffer exactly full
    s->overflow = true;
    *ret = (int64_t)s->off; *verdict = TBZ_OUTPUT_OVERFLOW;
    return TBZ_OK;
  }
  s->overflow = false;
  int32_t v = s->total.verdict;
  *verdict = v;
  if (v == TBZ_FINISHED) { s->finished = true; *ret = (int64_t)s->off; }
  else if (v == TBZ_INPUT_UNDERRUN) {
    s->underrun = true;
    bool zero = (s->format == TBZ_GZIP && s->total.where != TBZ_AT_BODY) ||
                (s->format == TBZ_ZLIB && s->total.where == TBZ_AT_HEADER);
    *ret = zero ? 0 : (int64_t)s->off;        // gzip.lisp:86,99,117 / zlib.lisp:111
  } else { s->error = v; *ret = -1; }
  return TBZ_OK;
}

// =============================================================================================
// gzip metadata and concatenated members.  Header rules as gzip.lisp:113-260.
// =============================================================================================
static uint32_t host_crc32(const uint8_t *p, uint64_t n) {
  uint32_t c = 0xffffffffu;
  for (uint64_t i = 0; i < n; i++) {
    c ^= p[i];
    for (int k = 0; k < 8; k++) c = (c >> 1) ^ ((c & 1) ? 0xedb88320u : 0);
  }
  return c ^ 0xffffffffu;
}

extern "C" int32_t tbz_gzip_header_parse(const uint8_t *in, uint64_t in_len, tbz_gzip_header *h) {
  if (!h || (in_len && !in)) return TBZ_E_ARG;
  memset(h, 0, sizeof *h);
  h->verdict = TBZ_INPUT_UNDERRUN;
  uint64_t p = 0;
  if (in_len < 2) return TBZ_OK;
  if (in[0] != 0x1f || in[1] != 0x8b) { h->verdict = TBZ_ERR_GZIP_MAGIC; return TBZ_OK; }
  if (in_len < 4) return TBZ_OK;
  if (in[2] != 8) { h->verdict = TBZ_ERR_GZIP_METHOD; return TBZ_OK; }
  const uint32_t flg = in[3];
  if (flg >> 5) { h->verdict = TBZ_ERR_GZIP_RESERVED; return TBZ_OK; }
  h->flags = flg & 31u;
  if (in_len < 8) return TBZ_OK;
  h->mtime = in[4] | ((uint32_t)in[5] << 8) | ((uint32_t)in[6] << 16) | ((uint32_t)in[7] << 24);
  if (in_len < 10) return TBZ_OK;
  h->xfl = in[8]; h->os = in[9];
  p = 10;
  if (flg & TBZ_GZ_EXTRA) {
    if (in_len - p < 2) return TBZ_OK;
    const uint64_t xlen = in[p] | ((uint64_t)in[p + 1] << 8);
    p += 2;
    if (in_len - p < xlen) return TBZ_OK;
    h->extra_off = p; h->extra_len = xlen;
    p += xlen;
  }
  if (flg & TBZ_GZ_NAME) {
    h->name_off = p;
    while (p < in_len && in[p]) p++;
    if (p >= in_len) return TBZ_OK;
    h->name_len = p - h->name_off;
    p++;
  }
  if (flg & TBZ_GZ_COMMENT) {
    h->comment_off = p;
    while (p < in_len && in[p]) p++;
    if (p >= in_len) return TBZ_OK;
    h->comment_len = p - h->comment_off;
    p++;
  }
  if (flg & TBZ_GZ_HCRC) {
    if (in_len - p < 2) return TBZ_OK;
    h->header_crc = in[p] | ((uint32_t)in[p + 1] << 8);
    if ((host_crc32(in, p) & 0xffffu) != h->header_crc) { h->verdict = TBZ_ERR_GZIP_HCRC; return TBZ_OK; }   // gzip.lisp:247-255
    p += 2;
  }
  h->header_len = p;
  h->verdict = TBZ_FINISHED;
  return TBZ_OK;
}

extern "C" int32_t tbz_inflate_gzip_members(tbz_ctx *ctx, const uint8_t *in, uint64_t in_len, uint8_t *out,
                                            uint64_t out_cap, tbz_result *r, uint64_t max_members,
                                            uint64_t *n_members, uint64_t *in_used) {
  if (!ctx || !r || !n_members || (in_len && !in) || (out_cap && !out)) return fail(ctx, TBZ_E_ARG, "tbz_inflate_gzip_members: bad argument");
  // the input goes to the device once, the members are decoded there one after the other (a member's length is only
  // known once it is decoded), and what they produced comes back with one copy
  CK(ctx, cudaSetDevice(ctx->device));
  void *d_in = nullptr, *d_out = nullptr;
  int32_t rc = dev_alloc(ctx, in_len + 16, &d_in);
  if (!rc) rc = dev_alloc(ctx, out_cap + 16, &d_out);
  if (rc) { dev_release(ctx, d_in); return rc; }
  uint64_t ip = 0, op = 0, k = 0, produced = 0;
  if (in_len) {
    cudaError_t e = cudaMemcpyAsync(d_in, in, in_len, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) rc = fail(ctx, TBZ_E_CUDA, "H2D", e);
  }
  while (!rc && k < max_members && ip < in_len) {
    tbz_member m{(const uint8_t *)d_in + ip, in_len - ip, (uint8_t *)d_out + op, out_cap - op};
    rc = tbz_inflate_batch(ctx, TBZ_GZIP, &m, 1, &r[k], TBZ_FLAG_DEVICE_PTRS, nullptr);
    if (rc != TBZ_OK) break;
    k++;
    produced = op + std::min<uint64_t>(r[k - 1].out_len, out_cap - op);   // (a member that did not finish leaves what it produced)
    if (r[k - 1].verdict != TBZ_FINISHED) break;
    ip += r[k - 1].in_used;
    op += r[k - 1].out_len;
  }
  if (!rc && produced) {
    cudaError_t e = cudaMemcpyAsync(out, d_out, produced, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) rc = fail(ctx, TBZ_E_CUDA, "D2H", e);
  }
  dev_release(ctx, d_in); dev_release(ctx, d_out);
  if (rc) return rc;
  *n_members = k;
  if (in_used) *in_used = ip;
  return TBZ_OK;
}
