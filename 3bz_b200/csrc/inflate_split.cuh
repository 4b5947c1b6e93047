// inflate_split.cuh — speculative split decode of ONE large member across the whole GPU
// (pugz / rapidgzip style; the reference has no equivalent: it decodes a stream strictly in order,
// deflate.lisp:92-730).
//
//   K0 k_split_find      the compressed body is cut at fixed offsets; a warp per chunk searches for the
//                        first dynamic-block start inside its chunk (find_block_start)
//   K1 k_split_decode    a warp per chunk decodes blocks from its start until it lands EXACTLY on a later
//                        candidate start (candidates it runs past were false positives of the search and
//                        are skipped): token slabs, output size and the bit it landed on.  The host follows
//                        the chain of landings from chunk 0 (chunks nobody lands on are dropped) and
//                        prefix-sums the output sizes; one pass, no re-decoding
//   K2 k_split_resolve   a CTA per chunk resolves its tokens with 16-bit symbols: bytes, or markers
//                        "byte i of the 32 KiB before this chunk" for what it cannot know yet
//   K3 k_tail_*          the last 32 KiB of output up to the end of every chunk become final by a
//                        parallel scan: per chunk a 32 Ki-entry map "window after the chunk, in terms
//                        of the window before it"; maps compose associatively (Hillis-Steele over the
//                        chunks, log2(chunks) rounds, every round fully parallel).  k_split_tails is the
//                        sequential form of the same step (one CTA walks the chunks), kept for comparison
//   K4 k_split_translate every other symbol becomes a byte, all chunks in parallel
//   K5 checksum          gzip: tbzcrc::k_span_crc (CRC-32 of 1 MiB pieces); zlib: k_split_adler (Adler-32 sums per
//                        64 KiB segment); the host combines the partials
#pragma once
#include "tbz_device.cuh"
#include "inflate_decode.cuh"
#include "inflate_resolve.cuh"

namespace tbzsplit {

constexpr uint64_t NONE64 = ~0ull;

struct Chunk {                           // host fills start/stop, the kernels fill the rest
  uint64_t start_bit;                    // absolute bit (from the member's 4-byte aligned base) of a block start
  uint64_t stop_bit;                     // decode until a block ends at or beyond this bit (NONE64: to the final block)
  uint64_t out_off;                      // absolute output offset of the chunk
  uint64_t land_bit;                     // K1: absolute bit after the chunk's last block
  tbzfast::P1Rec rec;                    // K1
  uint32_t ok;                           // K2
  uint32_t pad;
};

__device__ __forceinline__ tbzfast::In chunk_input(const uint32_t *words, uint64_t end_bit, uint64_t from_bit, uint32_t &rel) {
  tbzfast::In in;
  const uint64_t bw = from_bit >> 5;
  in.w = words + bw;
  rel = (uint32_t)(from_bit - (bw << 5));
  in.pos0 = rel;
  uint64_t e = end_bit - (bw << 5);
  if (e > 0xf0000000ull) e = 0xf0000000ull;
  in.end = (uint32_t)e;
  in.nwords = (in.end + 31) >> 5;
  return in;
}

// found[c] = first dynamic-block start in [c * chunk_bits, (c + 1) * chunk_bits) + body_bit, c >= 1.
// The chunk's bits are searched in pieces of FPIECE bits, dealt round-robin to FSUB warps, so the
// warps of a chunk advance through it together; the earliest hit wins (atomicMin), and a warp stops
// once a hit below its next piece is known.
constexpr uint32_t FSUB = 8, FPIECE = 4096;
__global__ void __launch_bounds__(tbzfast::NT)
k_split_find(const uint32_t *words, uint64_t end_bit, uint64_t body_bit, uint64_t chunk_bits, uint32_t nchunks, uint64_t *found) {
  TBZ_DYN_SMEM(smem_raw);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  tbzfast::WSmem &sm = reinterpret_cast<tbzfast::WSmem *>(smem_raw)[warp];
  const uint32_t gw = blockIdx.x * tbzfast::WPC + warp;
  const uint32_t c = gw / FSUB + 1, sub = gw % FSUB;
  if (c >= nchunks) return;
  const uint64_t cfrom = body_bit + chunk_bits * c;
  uint64_t cto = cfrom + chunk_bits;
  if (cto > end_bit) cto = end_bit;
  for (uint64_t from = cfrom + (uint64_t)FPIECE * sub; from < cto; from += (uint64_t)FPIECE * FSUB) {
    if (*reinterpret_cast<volatile unsigned long long *>(&found[c]) < from) break;   // an earlier hit exists
    uint64_t pe = from + FPIECE;
    if (pe > cto) pe = cto;
    uint32_t rel;
    const tbzfast::In in = chunk_input(words, end_bit, from, rel);
    const uint32_t r = tbzfast::find_block_start(in, rel, rel + (uint32_t)(pe - from), sm, lane);
    if (r != 0xffffffffu) {
      if (lane == 0) atomicMin(reinterpret_cast<unsigned long long *>(&found[c]), (unsigned long long)(from + (r - rel)));
      break;
    }
  }
}

// todo[i]: index of a chunk to decode.  cands: the ncands candidate starts in ascending order;
// chunks[c].pad = position of the chunk's own start in cands.
__global__ void __launch_bounds__(tbzfast::NT)
k_split_decode(const uint32_t *words, uint64_t end_bit, Chunk *chunks, const uint32_t *todo, uint32_t ntodo,
               uint32_t *slabs, uint32_t nslabs, uint32_t *counters, const unsigned long long *cands, uint32_t ncands) {
  TBZ_DYN_SMEM(smem_raw);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  tbzfast::WSmem &sm = reinterpret_cast<tbzfast::WSmem *>(smem_raw)[warp];
  for (;;) {
    uint32_t i = 0;
    if (lane == 0) i = atomicAdd(&counters[0], 1u);
    i = __shfl_sync(TBZ_FULL, i, 0);
    if (i >= ntodo) break;
    Chunk &ch = chunks[todo[i]];
    uint32_t rel;
    const tbzfast::In in = chunk_input(words, end_bit, ch.start_bit, rel);
    const uint64_t base = ch.start_bit - rel;
    tbzfast::StopList sl;
    sl.starts = cands; sl.n = ncands; sl.next = ch.pad + 1u; sl.base = base;
    const long long t0 = clock64();
    const bool ok = tbzfast::decode_blocks(in, rel, sl.rel(sl.next), 0xffffffffull, ch.rec, sm, slabs, nslabs, &counters[2], lane, &sl);
    __syncwarp();
    if (lane == 0) {
      if (!ok) ch.rec.status = 0;
      ch.land_bit = ok ? base + ch.rec.end_pos : NONE64;
      ch.pad = (uint32_t)((clock64() - t0) >> 10);      // (TBZ_KTIME: the chunk's decode time in 1 024-cycle units; the candidate index is used up)
    }
  }
}

typedef tbzres::SmemT<uint16_t> SymSmem;

__global__ void __launch_bounds__(tbzres::NT)
k_split_resolve(Chunk *chunks, uint32_t nchunks, const uint32_t *slabs, uint16_t *sym, uint32_t *counters) {
  TBZ_DYN_SMEM(smem_raw);
  SymSmem &sm = *reinterpret_cast<SymSmem *>(smem_raw);
  const int tid = threadIdx.x;
  for (;;) {
    __syncthreads();
    if (tid == 0) sm.member = atomicAdd(&counters[3], 1u);
    __syncthreads();
    const uint32_t k = sm.member;
    if (k >= nchunks) break;
    Chunk &ch = chunks[k];
    const uint32_t off = (uint32_t)ch.out_off;
    // what precedes the chunk is unknown: position a of the last 32 KiB is the marker for itself
    for (uint32_t i = tid; i < tbzres::HIST; i += tbzres::NT)
      sm.hist[(off - tbzres::HIST + i) & tbzres::HMASK] = (uint16_t)(tbzres::SYM_MARK | i);
    tbzres::RState rs;
    rs.pos = off; rs.flushed = off; rs.acc_a = 0; rs.acc_w = 0; rs.carry_len = 0; rs.carry_tok = 0;
    __syncthreads();
    const bool ok = tbzres::resolve_stream<uint16_t>(sym, TBZ_DEFLATE, ch.rec, slabs, rs, sm, tid);
    if (tid == 0) ch.ok = ok && rs.pos == off + ch.rec.out_len;
  }
}

// the last 32 KiB of every chunk, in order: the only sequential step.  One CTA of 1024 threads, 32
// symbols per thread.  The 32 KiB that precede the current chunk live in a shared-memory ring over
// absolute offsets (a marker costs a shared-memory load), and the tail symbols of chunk k+1 are
// fetched from HBM into registers while chunk k is resolved out of a shared-memory staging buffer,
// so no iteration waits for a memory round trip.
struct TailSmem { uint16_t stage[2][32768]; uint8_t ring[32768]; };

__device__ __forceinline__ void tail_range(const Chunk *chunks, uint32_t nchunks, uint32_t k, uint64_t &start, uint64_t &lo, uint64_t &end) {
  start = 0; lo = 0; end = 0;
  if (k >= nchunks) return;
  start = chunks[k].out_off;
  end = start + chunks[k].rec.out_len;
  lo = end > 32768 ? end - 32768 : 0;
  if (lo < start) lo = start;
}

__global__ void __launch_bounds__(1024)
k_split_tails(const Chunk *__restrict__ chunks, uint32_t nchunks, const uint16_t *__restrict__ sym, uint8_t *__restrict__ out) {
  TBZ_DYN_SMEM(smem_raw);
  TailSmem &sm = *reinterpret_cast<TailSmem *>(smem_raw);
  const uint32_t tid = threadIdx.x;
  uint64_t start, lo, end;
  tail_range(chunks, nchunks, 0, start, lo, end);
  for (int j = 0; j < 32; j++) {
    const uint64_t a = lo + tid + 1024u * j;
    sm.stage[0][tid + 1024u * j] = a < end ? sym[a] : (uint16_t)0;
  }
  __syncthreads();
  for (uint32_t k = 0; k < nchunks; k++) {
    uint64_t nstart, nlo, nend;
    tail_range(chunks, nchunks, k + 1, nstart, nlo, nend);
    uint16_t nx[32];                                   // next chunk's tail symbols: in flight during this chunk
#pragma unroll
    for (int j = 0; j < 32; j++) {
      const uint64_t a = nlo + tid + 1024u * j;
      nx[j] = a < nend ? sym[a] : (uint16_t)0;
    }
    const uint16_t *cur = sm.stage[k & 1];
    uint32_t pk[8];
#pragma unroll
    for (int j = 0; j < 32; j++) {
      uint32_t s = cur[tid + 1024u * j];
      if (s & tbzres::SYM_MARK) s = sm.ring[(start - 32768 + (s & 0x7fffu)) & 32767u];
      if ((j & 3) == 0) pk[j >> 2] = s; else pk[j >> 2] |= s << (8 * (j & 3));
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 32; j++) {
      const uint64_t a = lo + tid + 1024u * j;
      const uint8_t bv = (uint8_t)(pk[j >> 2] >> (8 * (j & 3)));
      if (a < end) { sm.ring[a & 32767u] = bv; out[a] = bv; }
    }
    uint16_t *nxt = sm.stage[(k + 1) & 1];
#pragma unroll
    for (int j = 0; j < 32; j++) nxt[tid + 1024u * j] = nx[j];
    __syncthreads();
    start = nstart; lo = nlo; end = nend;
  }
}

// ---- K3 as a scan --------------------------------------------------------------------------------
// map[v][j], j < 32768: the byte at absolute output offset end_v - 32768 + j as a symbol relative to
// the 32 KiB window that precedes chunk v: a byte, or SYM_MARK | i = "entry i of that window".
constexpr uint32_t TAILW = 32768;
__global__ void __launch_bounds__(256)
k_tail_init(const Chunk *__restrict__ chunks, const uint16_t *__restrict__ sym, uint16_t *__restrict__ map) {
  const uint32_t v = blockIdx.x;
  const long long start = (long long)chunks[v].out_off, len = (long long)chunks[v].rec.out_len, end = start + len;
  uint16_t *mv = map + (size_t)v * TAILW;
  for (uint32_t j = blockIdx.y * 1024u + threadIdx.x; j < blockIdx.y * 1024u + 1024u; j += 256u) {
    const long long a = end - (long long)TAILW + j;
    uint32_t s = 0;
    if (a >= start) s = sym[a];
    else if (a >= 0) s = tbzres::SYM_MARK | (uint32_t)(len + j);   // the chunk is shorter than the window: an older byte
    mv[j] = (uint16_t)s;
  }
}
// out[v] = in[v] o in[v - stride]: markers of map v are looked up in the map `stride` chunks back.  Eight entries per
// thread (one 16-byte load and store; the lookups are 2-byte gathers in a 64 KB map: L1 / L2).
__global__ void __launch_bounds__(256)
k_tail_compose(const uint16_t *__restrict__ in, uint16_t *__restrict__ outm, uint32_t stride) {
  const uint32_t v = blockIdx.x;
  const uint16_t *mv = in + (size_t)v * TAILW;
  uint16_t *ov = outm + (size_t)v * TAILW;
  const uint16_t *pv = v >= stride ? in + (size_t)(v - stride) * TAILW : nullptr;
  for (uint32_t j = (blockIdx.y * 256u + threadIdx.x) * 8u; j < TAILW; j += gridDim.y * 256u * 8u) {
    uint4 q = *reinterpret_cast<const uint4 *>(mv + j);
    if (pv && ((q.x | q.y | q.z | q.w) & (tbzres::SYM_MARK | (tbzres::SYM_MARK << 16)))) {
      uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
      for (int k = 0; k < 4; k++) {
        uint32_t lo = w[k] & 0xffffu, hi = w[k] >> 16;
        if (lo & tbzres::SYM_MARK) lo = pv[lo & 0x7fffu];
        if (hi & tbzres::SYM_MARK) hi = pv[hi & 0x7fffu];
        w[k] = lo | (hi << 16);
      }
      q = make_uint4(w[0], w[1], w[2], w[3]);
    }
    *reinterpret_cast<uint4 *>(ov + j) = q;
  }
}
// the final windows go to the output: chunk v owns [max(end - 32768, start), end)
__global__ void __launch_bounds__(256)
k_tail_write(const Chunk *__restrict__ chunks, const uint16_t *__restrict__ map, uint8_t *__restrict__ out) {
  const uint32_t v = blockIdx.x;
  const long long start = (long long)chunks[v].out_off, end = start + (long long)chunks[v].rec.out_len;
  const uint16_t *mv = map + (size_t)v * TAILW;
  for (uint32_t j = blockIdx.y * 1024u + threadIdx.x; j < blockIdx.y * 1024u + 1024u; j += 256u) {
    const long long a = end - (long long)TAILW + j;
    if (a >= start) out[a] = (uint8_t)mv[j];
  }
}

// every symbol becomes a byte, sixteen per thread and step (two 16-byte loads in flight, one 16-byte store).  The
// tails are final already and come out the same again: a marker's byte is read through L1 (the only concurrent
// writes to those addresses store the value they hold).  offs[k] = output offset of chunk k, offs[nchunks] = total.
__global__ void __launch_bounds__(256)
k_split_translate(const uint64_t *__restrict__ offs, uint32_t nchunks, const uint16_t *__restrict__ sym, uint8_t *out, uint64_t total) {
  const uint64_t nunits = (total + 15) / 16;
  // every block takes one contiguous range of units, so a thread stays inside one chunk for many
  // steps and the chunk lookup is two cached loads instead of a binary search
  const uint64_t per_block = (nunits + gridDim.x - 1) / gridDim.x;
  const uint64_t u_end = per_block * (blockIdx.x + 1) < nunits ? per_block * (blockIdx.x + 1) : nunits;
  const bool vec = (((uintptr_t)out) & 15) == 0;
  uint32_t k = 0xffffffffu;
  for (uint64_t u = per_block * blockIdx.x + threadIdx.x; u < u_end; u += blockDim.x) {
    const uint64_t a0 = u * 16;
    if (k == 0xffffffffu || a0 < offs[k] || a0 >= offs[k + 1]) {
      k = 0;                                             // last chunk with offs[k] <= a0
      for (uint32_t stp = 1u << 15; stp; stp >>= 1)
        if (k + stp < nchunks && offs[k + stp] <= a0) k += stp;
    }
    uint64_t start = offs[k], next = offs[k + 1];
    if (vec && a0 + 16 <= total && a0 + 16 <= next) {
      const uint4 v0 = __ldcs(reinterpret_cast<const uint4 *>(sym + a0)), v1 = __ldcs(reinterpret_cast<const uint4 *>(sym + a0 + 8));
      uint32_t s[16] = {v0.x & 0xffffu, v0.x >> 16, v0.y & 0xffffu, v0.y >> 16, v0.z & 0xffffu, v0.z >> 16, v0.w & 0xffffu, v0.w >> 16,
                        v1.x & 0xffffu, v1.x >> 16, v1.y & 0xffffu, v1.y >> 16, v1.z & 0xffffu, v1.z >> 16, v1.w & 0xffffu, v1.w >> 16};
      const uint8_t *const win = out + (start - 32768);
      if ((v0.x | v0.y | v0.z | v0.w | v1.x | v1.y | v1.z | v1.w) & (tbzres::SYM_MARK | (tbzres::SYM_MARK << 16))) {
#pragma unroll
        for (int j = 0; j < 16; j++)
          if (s[j] & tbzres::SYM_MARK) s[j] = win[s[j] & 0x7fffu];
      }
      uint4 o;
      o.x = s[0] | (s[1] << 8) | (s[2] << 16) | (s[3] << 24);
      o.y = s[4] | (s[5] << 8) | (s[6] << 16) | (s[7] << 24);
      o.z = s[8] | (s[9] << 8) | (s[10] << 16) | (s[11] << 24);
      o.w = s[12] | (s[13] << 8) | (s[14] << 16) | (s[15] << 24);
      *reinterpret_cast<uint4 *>(out + a0) = o;
    } else {
      for (uint64_t a = a0; a < a0 + 16 && a < total; a++) {
        while (a >= next) { k++; start = next; next = offs[k + 1]; }
        const uint32_t sv = sym[a];
        out[a] = (sv & tbzres::SYM_MARK) ? *reinterpret_cast<const volatile uint8_t *>(out + (start - 32768 + (sv & 0x7fffu))) : (uint8_t)sv;
      }
    }
  }
}

// Adler-32 partials of the output (zlib members; gzip goes through tbzcrc::k_span_crc): one WARP per 64 KiB segment,
// coalesced 16-byte loads, dp4a sums.  parts[2 seg] = sum of the segment's bytes, parts[2 seg + 1] = sum (m - j) d_j
// (m = the segment's length, j = offset in it), both mod 65521; the host chains the segments
// (s2 += m s1 + parts[1]; s1 += parts[0]).  Round 1 ran one THREAD per 4 KiB segment (every lane its own 4 KiB stride).
constexpr uint32_t ASEG = 65536;
__global__ void __launch_bounds__(256)
k_split_adler(const uint8_t *__restrict__ out, uint64_t n, uint32_t *parts) {
  const int lane = threadIdx.x & 31;
  const uint64_t seg = (uint64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const uint64_t lo = seg * ASEG;
  if (lo >= n) return;
  const uint64_t hi = lo + ASEG < n ? lo + ASEG : n;
  const uint32_t m = (uint32_t)(hi - lo);
  unsigned long long a = 0, w = 0;
  uint64_t i = lo;
  if ((((uintptr_t)(out + lo)) & 15) == 0) {
    const uint64_t vend = lo + (m & ~15u);
    for (uint64_t p = lo + 16u * lane; p < vend; p += 512u) {
      const uint4 v = __ldcs(reinterpret_cast<const uint4 *>(out + p));
      uint32_t sd = __dp4a(v.x, 0x01010101u, 0u); sd = __dp4a(v.y, 0x01010101u, sd);
      sd = __dp4a(v.z, 0x01010101u, sd); sd = __dp4a(v.w, 0x01010101u, sd);
      uint32_t wj = __dp4a(v.x, 0x03020100u, 0u); wj = __dp4a(v.y, 0x07060504u, wj);
      wj = __dp4a(v.z, 0x0b0a0908u, wj); wj = __dp4a(v.w, 0x0f0e0d0cu, wj);
      a += sd;
      w += (unsigned long long)(m - (uint32_t)(p - lo)) * sd - wj;       // sum (m - j) d_j over the unit
    }
    i = vend;
  }
  for (uint64_t p = i + lane; p < hi; p += 32) {                         // what 16-byte loads cannot take
    const uint32_t d = out[p];
    a += d; w += (unsigned long long)(m - (uint32_t)(p - lo)) * d;
  }
  a %= TBZ_ADLER_MOD; w %= TBZ_ADLER_MOD;
#pragma unroll
  for (int sft = 16; sft; sft >>= 1) { a += __shfl_xor_sync(TBZ_FULL, a, sft); w += __shfl_xor_sync(TBZ_FULL, w, sft); }
  if (lane == 0) { parts[2 * seg] = (uint32_t)(a % TBZ_ADLER_MOD); parts[2 * seg + 1] = (uint32_t)(w % TBZ_ADLER_MOD); }
}

}  // namespace tbzsplit
