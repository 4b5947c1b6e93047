// cuda_emu.cpp — scheduler of the SIMT emulator (see cuda_emu.h).  Test infrastructure only.
#include "cuda_emu.h"

#include <mutex>

thread_local EmuIdx threadIdx, blockIdx, blockDim, gridDim;
int emu_os_threads = 4;

namespace emu {

thread_local Block *g_blk = nullptr;
static const size_t kStack = 256 * 1024;

void die(const char *what) {
  Block *b = g_blk;
  fprintf(stderr, "[cuda_emu] FATAL: %s (block %u thread %u)\n", what, b ? b->bidx.x : 0u, b ? b->cur : 0u);
  fflush(stderr);
  abort();
}

unsigned char *dyn_smem() { return g_blk->dyn.data(); }

static void release_barrier(Block *b) {
  b->bar_or_out[b->bar_gen & 1] = b->bar_or;
  b->bar_or = 0; b->bar_arrived = 0; b->bar_gen++; b->epoch++;
}

void block_barrier(int pred, int *or_out) {
  Block *b = g_blk;
  Warp &w = b->warps[b->cur >> 5];
  if (w.arrived) die("__syncthreads while lanes of the same warp wait in a warp collective");
  b->bar_or |= pred ? 1 : 0;
  b->bar_arrived++;
  const unsigned g = b->bar_gen;
  if (b->bar_arrived == b->bar_alive) release_barrier(b);
  else while (b->bar_gen == g) yield();
  if (or_out) *or_out = b->bar_or_out[g & 1];
}

static void fiber_entry() {
  Block *b = g_blk;
  b->body();
  const unsigned tid = b->cur;
  Warp &w = b->warps[tid >> 5];
  if (w.arrived && (w.mask >> (tid & 31) & 1)) die("a lane exited while its warp waits for it in a collective");
  w.alive &= ~(1u << (tid & 31));
  b->fibers[tid].done = true;
  b->bar_alive--;
  b->epoch++;
  if (b->bar_arrived && b->bar_arrived == b->bar_alive) release_barrier(b);
  swapcontext(&b->fibers[tid].ctx, &b->sched);
}

static void run_block(Block *b, unsigned bx, dim3 grid, dim3 block, size_t smem, const std::function<void()> &body) {
  g_blk = b;
  const unsigned n = block.x * block.y * block.z;
  b->nthreads = n;
  b->bidx = dim3(bx % grid.x, bx / grid.x, 0); b->bdim = block; b->gdim = grid;
  b->body = body;
  if (b->fibers.size() < n) b->fibers.resize(n);
  b->warps.assign((n + 31) / 32, Warp());
  b->dyn.assign(smem, 0xcd);       // exactly the bytes the launch asked for; shared memory is not zero-initialised on hardware either
  b->bar_arrived = 0; b->bar_alive = n; b->bar_gen = 0; b->bar_or = 0; b->epoch = 0;
  for (unsigned t = 0; t < n; t++) {
    Fiber &f = b->fibers[t];
    if (f.stack.empty()) f.stack.resize(kStack);
    f.done = false;
    getcontext(&f.ctx);
    f.ctx.uc_stack.ss_sp = f.stack.data();
    f.ctx.uc_stack.ss_size = f.stack.size();
    f.ctx.uc_link = nullptr;
    makecontext(&f.ctx, fiber_entry, 0);
    b->warps[t >> 5].alive |= 1u << (t & 31);
  }
  unsigned alive = n;
  while (alive) {
    const uint64_t e0 = b->epoch;
    for (unsigned t = 0; t < n; t++) {
      Fiber &f = b->fibers[t];
      if (f.done) continue;
      b->cur = t;
      threadIdx = EmuIdx{t % block.x, (t / block.x) % block.y, t / (block.x * block.y)};
      blockIdx = EmuIdx{bx % grid.x, bx / grid.x, 0};
      blockDim = EmuIdx{block.x, block.y, block.z};
      gridDim = EmuIdx{grid.x, grid.y, grid.z};
      swapcontext(&b->sched, &f.ctx);
      if (f.done) alive--;
    }
    if (alive && b->epoch == e0) {
      b->cur = 0;
      die("deadlock: a full scheduling pass made no progress (a collective or barrier that not every thread reaches)");
    }
  }
  g_blk = nullptr;
}

void launch_impl(dim3 grid, dim3 block, size_t smem, const std::function<void()> &body, int os_threads) {
  if (grid.z != 1) { fprintf(stderr, "[cuda_emu] only 1-D and 2-D grids\n"); abort(); }
  const unsigned nblocks = grid.x * grid.y;
  std::atomic<unsigned> next{0};
  auto worker = [&]() {
    Block *b = new Block();
    for (;;) {
      const unsigned bx = next.fetch_add(1);
      if (bx >= nblocks) break;
      run_block(b, bx, grid, block, smem, body);
    }
    delete b;
  };
  const int nt = (int)std::min<unsigned>((unsigned)std::max(1, os_threads), nblocks);
  if (nt <= 1) { worker(); return; }
  std::vector<std::thread> th;
  for (int i = 0; i < nt; i++) th.emplace_back(worker);
  for (auto &t : th) t.join();
}

}  // namespace emu
