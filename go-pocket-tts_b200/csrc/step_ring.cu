// step_ring.cu — the batch-1 AR frame with the weights streamed through a shared-memory ring
// (PTTS_FUSED_RING; same 46 phases and the same arithmetic as step_fused.cu).
//
// step_fused_kernel loads a phase's weight rows into registers, so the HBM stream stops at every
// phase boundary: per phase the costs ADD (bytes / ~3 TB/s + exchange + staging).  The weights
// do not depend on the activations, so here ONE producer warp per CTA walks the CTA's share of
// the whole frame's weights — every weight phase in execution order — and issues 1-D bulk copies
// (cp.async.bulk, completion by mbarrier transaction bytes: no registers, outside the L1tex FIFO)
// into a ring of 4 KB slots, limited only by the slots the consumers have released.  With 48
// slots per SM the stream runs ~28 MB (about half a transformer layer) ahead of the arithmetic:
// it keeps going while the consumer warps sit in a grid barrier, stage LayerNorm'd activations or
// run the attention phase, and the consumers find their rows already in shared memory.
//
// Geometry: one CTA per SM, 4 consumer warps + 1 producer warp.  CTA b owns the contiguous rows
// [N b / G, N (b + 1) / G) of every weight matrix (a contiguous block of memory); inside the
// block, rows go round-robin to the consumer warps, a row's 1024-wide k chunks stay with its
// warp.  Ring order = the block front to back (row, chunk); producer and consumers derive a
// segment's slot from the same running count.  Per-lane summation order, chunk order and the shuffle
// reduction are those of step_fused.cu's gemv_cols / gemv_pairs, so the two kernels produce
// bit-identical frames.  The consumers synchronise among themselves on named barrier 1 (the
// producer never joins): __syncthreads() in the shared helpers of attend.cuh is mapped to it.
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <type_traits>

namespace ptts {
constexpr int RT = 128; // consumer threads (= ATT_THREADS of attend_row)
__device__ __forceinline__ void ring_consumer_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }
} // namespace ptts
#define __syncthreads() ::ptts::ring_consumer_sync()
#include "attend.cuh"
#include "gemm_epi.cuh"
#include "step_fused.cuh"

namespace ptts {

namespace {

constexpr int RING_SEG_K = 1024;                 // floats per slot
constexpr int RING_SLOT_BYTES = RING_SEG_K * 4;  // 4 KB
constexpr int RING_MAX_SLOTS = 52;
constexpr int RING_WARPS = RT / 32;

__device__ __forceinline__ uint32_t s_u32(const void *p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ float ldcg1(const float *p) { return __ldcg(p); }
__device__ __forceinline__ float bias_at(const float *b, int n) { return b ? __ldg(b + n) : 0.0f; }
__device__ __forceinline__ unsigned long long gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define RSTAMP()                                                               \
  do {                                                                         \
    if (P.dbg && blockIdx.x == 0 && threadIdx.x == 0 && nstamp < 500)          \
      P.dbg[nstamp++] = gtime();                                               \
  } while (0)
__device__ __forceinline__ void l1_prefetch(const void *p, size_t bytes) {
  if (!p)
    return;
  const char *c = reinterpret_cast<const char *>(p);
  for (size_t off = (size_t)threadIdx.x * 128; off < bytes; off += (size_t)RT * 128)
    asm volatile("prefetch.global.L1 [%0];" ::"l"(c + off));
}

// ---- mbarrier helpers with bounded waits (a missing arrival raises the error flag, never hangs) ----
__device__ __forceinline__ void mb_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mb_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s_u32(bar)) : "memory");
}
__device__ __forceinline__ void mb_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mb_try(uint64_t *bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\t"
               "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
               "selp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok)
               : "r"(s_u32(bar)), "r"(parity)
               : "memory");
  return ok != 0;
}
// returns false (and raises the flag) when the wait gave up
__device__ __forceinline__ bool mb_wait(uint64_t *bar, uint32_t parity, int *err) {
  long spins = 0;
  while (!mb_try(bar, parity)) {
    if (++spins > 4000000L || ((spins & 1023) == 0 && *reinterpret_cast<volatile int *>(err) != 0)) {
      atomicExch(err, 1);
      return false;
    }
  }
  return true;
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s_u32(dst)),
               "l"(src), "r"(bytes), "r"(s_u32(bar))
               : "memory");
}

// The frame's weight phases in execution order (the adaLN projection as row pairs when its K is half
// a slot, see step_fused.cu gemv_pairs).
__device__ __forceinline__ bool ring_phase(const FusedParams &P, int idx, int lsd_steps, bool pair_ada, const float *&W,
                                           int &N, int &K) {
  if (idx < 4 * P.L) {
    const FusedLayer &ly = P.layer[idx >> 2];
    switch (idx & 3) {
    case 0: W = ly.in_proj; N = 3 * P.D; K = P.D; break;
    case 1: W = ly.out_proj; N = P.D; K = P.D; break;
    case 2: W = ly.l1; N = P.F; K = P.D; break;
    default: W = ly.l2; N = P.D; K = P.F; break;
    }
    return true;
  }
  idx -= 4 * P.L;
  if (idx == 0) {
    W = P.ce_w; N = P.C; K = P.D;
    return true;
  }
  idx -= 1;
  const int per = 2 + 2 * P.depth;
  if (idx >= lsd_steps * per)
    return false;
  const int j = idx % per;
  if (j == 0) {
    W = P.ada_w;
    N = pair_ada ? P.ada_n / 2 : P.ada_n;
    K = pair_ada ? 2 * P.C : P.C;
  } else if (j == per - 1) {
    W = P.fin_w; N = P.ldim; K = P.C;
  } else {
    const FusedRes &rb = P.rb[(j - 1) >> 1];
    if ((j - 1) & 1) {
      W = rb.mlp2w; N = P.C; K = P.hid;
    } else {
      W = rb.mlp0w; N = P.hid; K = P.C;
    }
  }
  return true;
}
__device__ __forceinline__ void cta_rows(int N, int &n0, int &nrows) {
  n0 = (int)(((long)N * blockIdx.x) / gridDim.x);
  nrows = (int)(((long)N * (blockIdx.x + 1)) / gridDim.x) - n0;
}

// ---- producer: one warp, the CTA's share of every weight phase, in ring order ----
// Segment i of a phase = (row i / kc, chunk i % kc) — row-major, i.e. the CTA's block of the matrix front to
// back.  Lane j issues the segments i = j (mod 32): a single thread's wait -> expect_tx -> copy chain is
// ~200 ns per 4 KB (20 GB/s per SM, measured: the first version of this kernel), 32 lanes overlap the chains.
__device__ void ring_produce(const FusedParams &P, int lsd_steps, bool pair_ada, float *ring, uint64_t *full,
                             uint64_t *empty, uint32_t S, int lane) {
  uint32_t base = 0;
  for (int idx = 0;; idx++) {
    const float *W;
    int N, K;
    if (!ring_phase(P, idx, lsd_steps, pair_ada, W, N, K))
      break;
    int n0, nrows;
    cta_rows(N, n0, nrows);
    const int kc = (K + RING_SEG_K - 1) / RING_SEG_K;
    const int T = nrows * kc;
    const float *Wb = W + (long)n0 * K;
    for (int i0 = 0; i0 < T; i0 += 32) {
      const int i = i0 + lane;
      const int row = i / kc, c = i - row * kc;
      const int kn = K - c * RING_SEG_K < RING_SEG_K ? K - c * RING_SEG_K : RING_SEG_K;
      const uint32_t q = base + (uint32_t)i, round = q / S, slot = q - round * S;
      // The lanes stay in lockstep (one uniform loop, each lane issues the moment ITS slot is free): a lane's
      // slot never depends on a segment of the same or a later batch, so no lane waits for a sibling — nothing
      // here relies on the scheduler being fair to diverged lanes.
      bool pending = i < T;
      long spins = 0;
      while (__any_sync(0xffffffffu, pending)) {
        if (pending && (round == 0 || mb_try(&empty[slot], (round - 1) & 1))) {
          mb_expect_tx(&full[slot], (uint32_t)kn * 4);
          bulk_g2s(ring + (size_t)slot * RING_SEG_K, Wb + (long)row * K + (long)c * RING_SEG_K, (uint32_t)kn * 4,
                   &full[slot]);
          pending = false;
        }
        if (++spins > 4000000L || ((spins & 1023) == 0 && *reinterpret_cast<volatile int *>(P.error) != 0)) {
          if (lane == 0)
            atomicExch(P.error, 1);
          return;
        }
      }
    }
    base += (uint32_t)T;
  }
}

// ---- consumers ----
struct Ring {
  const float *ring;
  uint64_t *full, *empty;
  int *err;
  uint32_t S, base; // slots; segments consumed by this CTA before the current phase
  bool dead;
};
__device__ __forceinline__ float dot4(const float4 &a, const float4 &w, float acc) {
  acc = fmaf(a.x, w.x, acc);
  acc = fmaf(a.y, w.y, acc);
  acc = fmaf(a.z, w.z, acc);
  acc = fmaf(a.w, w.w, acc);
  return acc;
}
// wait for segment q, return its slot
__device__ __forceinline__ uint32_t ring_acquire(Ring &r, uint32_t q) {
  const uint32_t round = q / r.S, slot = q - round * r.S;
  if (!r.dead && !mb_wait(&r.full[slot], round & 1, r.err))
    r.dead = true;
  return slot;
}
__device__ __forceinline__ void ring_release(Ring &r, uint32_t slot, int lane) {
  __syncwarp();
  if (lane == 0)
    mb_arrive(&r.empty[slot]);
}
// y[n] = W[n, :K] . a for the CTA's rows; a is in shared memory; epi(n, y) runs on lane 0 of the row's warp.
template <class Epi>
__device__ __forceinline__ void ring_gemv(Ring &r, int N, int K, const float *sA, int warp, int lane, Epi epi) {
  int n0, nrows;
  cta_rows(N, n0, nrows);
  const int kc = (K + RING_SEG_K - 1) / RING_SEG_K;
  for (int row = warp; row < nrows; row += RING_WARPS) {
    float colacc = 0.0f;
    for (int c = 0; c < kc; c++) {
      const uint32_t slot = ring_acquire(r, r.base + (uint32_t)(kc * row + c));
      const float *w = r.ring + (size_t)slot * RING_SEG_K;
      const int k0 = c * RING_SEG_K;
      float acc = 0.0f;
#pragma unroll
      for (int u = 0; u < 8; u++) {
        const int kk = lane * 4 + 128 * u;
        if (k0 + kk < K)
          acc = dot4(*reinterpret_cast<const float4 *>(sA + k0 + kk), *reinterpret_cast<const float4 *>(w + kk), acc);
      }
      ring_release(r, slot, lane);
      colacc += acc;
    }
    float v = colacc;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
      v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0)
      epi(n0 + row, v);
  }
  r.base += (uint32_t)(nrows * kc);
}
// the [NP, 1024] row-pair view of a [2 NP, 512] matrix: the two halves of a slot are two rows
template <class Epi>
__device__ __forceinline__ void ring_gemv_pairs(Ring &r, int NP, const float *sA2, int warp, int lane, Epi epi) {
  int n0, nrows;
  cta_rows(NP, n0, nrows);
  for (int row = warp; row < nrows; row += RING_WARPS) {
    const uint32_t slot = ring_acquire(r, r.base + (uint32_t)row);
    const float *w = r.ring + (size_t)slot * RING_SEG_K;
    float lo = 0.0f, hi = 0.0f;
#pragma unroll
    for (int u = 0; u < 4; u++)
      lo = dot4(*reinterpret_cast<const float4 *>(sA2 + lane * 4 + 128 * u),
                *reinterpret_cast<const float4 *>(w + lane * 4 + 128 * u), lo);
#pragma unroll
    for (int u = 4; u < 8; u++)
      hi = dot4(*reinterpret_cast<const float4 *>(sA2 + lane * 4 + 128 * u),
                *reinterpret_cast<const float4 *>(w + lane * 4 + 128 * u), hi);
    ring_release(r, slot, lane);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      lo += __shfl_xor_sync(0xffffffffu, lo, o);
      hi += __shfl_xor_sync(0xffffffffu, hi, o);
    }
    if (lane == 0) {
      const int m = n0 + row;
      epi(2 * m, lo);
      epi(2 * m + 1, hi);
    }
  }
  r.base += (uint32_t)nrows;
}
// the tiny matrices (K = latent dim, the EOS row) stay on plain loads: rows gw, gw + nw, ...
template <class Epi>
__device__ __forceinline__ void gemv_direct(const float *W, int N, int K, const float *sA, int gw, int nw, int lane,
                                            Epi epi) {
  const int kc = (K + RING_SEG_K - 1) / RING_SEG_K;
  for (int n = gw; n < N; n += nw) {
    float colacc = 0.0f;
    for (int c = 0; c < kc; c++) {
      float acc = 0.0f;
#pragma unroll
      for (int u = 0; u < 8; u++) {
        const int k = c * RING_SEG_K + lane * 4 + 128 * u;
        if (k < K)
          acc = dot4(*reinterpret_cast<const float4 *>(sA + k), __ldg(reinterpret_cast<const float4 *>(W + (long)n * K + k)),
                     acc);
      }
      colacc += acc;
    }
    float v = colacc;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
      v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0)
      epi(n, v);
  }
}

// ---- grid barrier among the consumers of all CTAs (step_fused.cu's striped counters) ----
constexpr int GB_NCTR = 16;
constexpr int GB_STRIDE = 32;
struct GridBar {
  unsigned *ctr;
  int *err;
  unsigned phase, mine;
  bool dead;
};
__device__ __forceinline__ GridBar make_bar(unsigned *ctr, int *err) {
  GridBar b;
  b.ctr = ctr;
  b.err = err;
  b.phase = 0;
  b.dead = false;
  const unsigned c = threadIdx.x & (GB_NCTR - 1);
  b.mine = (gridDim.x + GB_NCTR - 1 - c) / GB_NCTR;
  return b;
}
__device__ __forceinline__ void grid_arrive(GridBar &b) {
  b.phase++;
  ring_consumer_sync();
  if (threadIdx.x == 0) {
    __threadfence();
    asm volatile("red.relaxed.gpu.global.add.u32 [%0], 1;" ::"l"(b.ctr + GB_STRIDE * (blockIdx.x & (GB_NCTR - 1)))
                 : "memory");
  }
}
__device__ __forceinline__ void grid_wait(GridBar &b) {
  if (threadIdx.x < GB_NCTR && !b.dead) {
    const unsigned target = b.phase * b.mine;
    const unsigned *p = b.ctr + GB_STRIDE * threadIdx.x;
    long spins = 0;
    for (;;) {
      unsigned v;
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
      if (v >= target)
        break;
      if (++spins > 8000000L || ((spins & 255) == 0 && *reinterpret_cast<volatile int *>(b.err) != 0)) {
        atomicExch(b.err, 1);
        break;
      }
    }
  }
  ring_consumer_sync();
  if (!b.dead && *reinterpret_cast<volatile int *>(b.err) != 0)
    b.dead = true;
}

__device__ __forceinline__ float block_sum(float v, float *red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
    v += __shfl_xor_sync(0xffffffffu, v, o);
  ring_consumer_sync();
  if ((threadIdx.x & 31) == 0)
    red[threadIdx.x >> 5] = v;
  ring_consumer_sync();
  return (red[0] + red[1]) + (red[2] + red[3]);
}
__device__ __forceinline__ void stage_plain(float *sA, const float *src, int K) {
  for (int i = threadIdx.x; i < K; i += RT)
    sA[i] = ldcg1(src + i);
  ring_consumer_sync();
}
// sA = (LN(src) * w + b) [* (1 + scale) + shift]  (linear.go:265-329, tensor_util.go:175-193)
__device__ __forceinline__ void stage_ln(float *sA, float *red, const float *src, int K, const float *w, const float *b,
                                         float eps, const float *scale, const float *shift) {
  float s = 0.0f;
  for (int i = threadIdx.x; i < K; i += RT) {
    const float v = ldcg1(src + i);
    sA[i] = v;
    s += v;
  }
  const float mean = block_sum(s, red) / (float)K;
  float q = 0.0f;
  for (int i = threadIdx.x; i < K; i += RT) {
    const float d = sA[i] - mean;
    q += d * d;
  }
  const float inv = 1.0f / sqrtf(block_sum(q, red) / (float)K + eps);
  for (int i = threadIdx.x; i < K; i += RT) {
    float t = (sA[i] - mean) * inv;
    if (w)
      t *= __ldg(w + i);
    if (b)
      t += __ldg(b + i);
    if (scale)
      t = t * (ldcg1(scale + i) + 1.0f) + ldcg1(shift + i);
    sA[i] = t;
  }
  ring_consumer_sync();
}

constexpr int RING_BLOCK = RT + 32; // + the producer warp

} // namespace

__global__ void __launch_bounds__(RING_BLOCK, 1)
    step_ring_kernel(const FusedParams *__restrict__ Pp, const int32_t *row_slot, const int32_t *row_pos,
                     const int64_t *off_a, const int64_t *off_b, const int64_t *off_c, int gather_frames,
                     int gather_noise, const float *tvec, int lsd_steps, int S) {
  extern __shared__ __align__(128) float dsm[]; // [S][1024] weight ring | activation vector / attention scores
  __shared__ __align__(8) uint64_t full_bar[RING_MAX_SLOTS], empty_bar[RING_MAX_SLOTS];
  __shared__ float red[RT / 32];
  __shared__ float4 pv[8 * 16];
  __shared__ __align__(16) float q_s[HD];
  const FusedParams &P = *Pp;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int D = P.D, F = P.F, C = P.C, ldim = P.ldim;
  const bool pair_ada = 2 * C == RING_SEG_K && (P.ada_n & 1) == 0;
  if (threadIdx.x == 0) {
    for (int i = 0; i < S; i++) {
      mb_init(&full_bar[i], 1);
      mb_init(&empty_bar[i], 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("bar.sync 0;" ::: "memory"); // the only CTA-wide barrier: the producer leaves after it
  if (warp == RING_WARPS) {
    ring_produce(P, lsd_steps, pair_ada, dsm, full_bar, empty_bar, (uint32_t)S, lane);
    return;
  }
  float *sA = dsm + (size_t)S * RING_SEG_K;
  Ring ring{dsm, full_bar, empty_bar, P.error, (uint32_t)S, 0u, false};
  const int gw = blockIdx.x * RING_WARPS + warp, nw = gridDim.x * RING_WARPS;
  GridBar bar = make_bar(P.barrier, P.error);
  const int slot = row_slot[0], pos = row_pos[0];
  int nstamp = 0;
  RSTAMP();
#define PHASE_BEGIN_LN(LW, LB, KLN)                                            \
  do {                                                                         \
    grid_arrive(bar);                                                          \
    l1_prefetch(LW, (size_t)(KLN) * 4);                                        \
    l1_prefetch(LB, (size_t)(KLN) * 4);                                        \
    grid_wait(bar);                                                            \
    RSTAMP();                                                                  \
  } while (0)
#define PHASE_BEGIN() PHASE_BEGIN_LN(nullptr, nullptr, 0)

  // ---- phase 0: x = input_linear(nan_to_bos(frame)) ----
  {
    const float *frame = gather_frames ? P.hist + off_a[0] : P.fh_frames;
    if (threadIdx.x < ldim) {
      const float v = ldcg1(frame + threadIdx.x);
      const float u = isnan(v) ? __ldg(P.bos + threadIdx.x) : v;
      sA[threadIdx.x] = u;
      if (blockIdx.x == 0)
        P.fh_in[threadIdx.x] = u;
    }
    if (blockIdx.x == 0 && threadIdx.x < ldim) {
      const float *nz = gather_noise ? P.noise_pool + off_c[0] : P.fh_noise;
      P.fh_cur[threadIdx.x] = ldcg1(nz + threadIdx.x);
    }
    ring_consumer_sync();
    gemv_direct(P.in_w, D, ldim, sA, gw, nw, lane, [&](int n, float v) { P.x[n] = v + bias_at(P.in_b, n); });
  }
  for (int l = 0; l < P.L; l++) {
    const FusedLayer &ly = P.layer[l];
    // ---- qkv = in_proj(LN1(x)) ----
    PHASE_BEGIN_LN(ly.n1w, ly.n1b, D);
    stage_ln(sA, red, P.x, D, ly.n1w, ly.n1b, ly.eps1, nullptr, nullptr);
    RSTAMP();
    ring_gemv(ring, 3 * D, D, sA, warp, lane, [&](int n, float v) { P.qkv[n] = v; });
    // ---- RoPE, KV append, attention: one CTA per head (the ring keeps filling meanwhile) ----
    PHASE_BEGIN();
    if ((int)blockIdx.x < P.H) {
      const int h = blockIdx.x, half = HD / 2;
      const KvGeom &kv = P.kv;
      const int PT = kv.page_tokens;
      const long page_stride = (long)kv.n_layers * 2 * kv.n_heads * PT * HD;
      const int32_t *pt = kv.pagetab + (long)slot * kv.max_pages;
      __half *page = static_cast<__half *>(kv.pool) + (long)pt[pos / PT] * page_stride;
      const int tok = pos % PT;
      const long koff = ((long)(l * 2 + 0) * kv.n_heads + h) * PT * HD;
      const long voff = ((long)(l * 2 + 1) * kv.n_heads + h) * PT * HD;
      if (threadIdx.x < half) { // interleaved pairs (rope.go:81-105)
        const int j = threadIdx.x;
        const float c = __ldg(P.rope_cos + (long)pos * half + j), s = __ldg(P.rope_sin + (long)pos * half + j);
        const float qa = ldcg1(P.qkv + h * HD + 2 * j), qb = ldcg1(P.qkv + h * HD + 2 * j + 1);
        q_s[2 * j] = rope_even(qa, qb, c, s);
        q_s[2 * j + 1] = rope_odd(qa, qb, c, s);
        const float ka = ldcg1(P.qkv + D + h * HD + 2 * j), kb = ldcg1(P.qkv + D + h * HD + 2 * j + 1);
        __half *dst = page + koff + (long)tok * HD + 2 * j;
        kv_store(dst, rope_even(ka, kb, c, s));
        kv_store(dst + 1, rope_odd(ka, kb, c, s));
      } else if (threadIdx.x < half + HD) {
        const int d = threadIdx.x - half;
        kv_store(page + voff + (long)tok * HD + d, ldcg1(P.qkv + 2 * D + h * HD + d));
      }
      ring_consumer_sync();
      const __half *pool = static_cast<const __half *>(kv.pool);
      auto keyp = [&](int j) { return pool + (long)pt[j / PT] * page_stride + koff + (long)(j % PT) * HD; };
      auto valp = [&](int j) { return pool + (long)pt[j / PT] * page_stride + voff + (long)(j % PT) * HD; };
      if (PT == 16)
        attend_row_flash16(q_s, pos + 1, pool + koff, pool + voff, pt, page_stride, P.att + h * HD, sA, SplitOut(), 0, 0);
      else
        attend_row(q_s, pos + 1, keyp, valp, P.att + h * HD, sA, red, pv, SplitOut(), 0, 0);
    }
    // ---- x += out_proj(att) ----
    PHASE_BEGIN();
    stage_plain(sA, P.att, D);
    RSTAMP();
    ring_gemv(ring, D, D, sA, warp, lane, [&](int n, float v) { P.x[n] = ldcg1(P.x + n) + v; });
    // ---- ff = gelu(linear1(LN2(x))) ----
    PHASE_BEGIN_LN(ly.n2w, ly.n2b, D);
    stage_ln(sA, red, P.x, D, ly.n2w, ly.n2b, ly.eps2, nullptr, nullptr);
    RSTAMP();
    ring_gemv(ring, F, D, sA, warp, lane, [&](int n, float v) { P.ff[n] = act_apply(v, ACT_GELU); });
    // ---- x += linear2(ff) ----
    PHASE_BEGIN();
    stage_plain(sA, P.ff, F);
    RSTAMP();
    ring_gemv(ring, D, F, sA, warp, lane, [&](int n, float v) { P.x[n] = ldcg1(P.x + n) + v; });
  }
  // ---- last = out_norm(x); eos = last . w_eos + b; cproj = cond_embed(last) ----
  PHASE_BEGIN_LN(P.on_w, P.on_b, D);
  stage_ln(sA, red, P.x, D, P.on_w, P.on_b, P.on_eps, nullptr, nullptr);
  RSTAMP();
  if (blockIdx.x == 0)
    for (int i = threadIdx.x; i < D; i += RT)
      P.fh_last[i] = sA[i];
  ring_gemv(ring, C, D, sA, warp, lane, [&](int n, float v) { P.fh_cproj[n] = v + bias_at(P.ce_b, n); });
  if (gw == nw - 1) // the EOS logit is one more row, on the last warp
    gemv_direct(P.eos_w, 1, D, sA, 0, 1, lane, [&](int, float v) { P.fh_eos[0] = v + P.eos_b; });
  const float inv_steps = 1.0f / (float)lsd_steps;
  for (int it = 0; it < lsd_steps; it++) {
    // ---- ada = adaLN_all(silu(cproj + t_emb)); fx = input_proj(cur) ----
    float *sCur = sA + (pair_ada ? 2 * C : C);
    PHASE_BEGIN();
    for (int i = threadIdx.x; i < C; i += RT) {
      float y = ldcg1(P.fh_cproj + i) + __ldg(tvec + (long)it * C + i);
      y = y / (1.0f + expf(-y));
      sA[i] = y;
      if (pair_ada)
        sA[C + i] = y;
    }
    for (int i = threadIdx.x; i < ldim; i += RT)
      sCur[i] = ldcg1(P.fh_cur + i);
    ring_consumer_sync();
    RSTAMP();
    if (pair_ada)
      ring_gemv_pairs(ring, P.ada_n / 2, sA, warp, lane, [&](int n, float v) { P.fh_ada[n] = v + bias_at(P.ada_b, n); });
    else
      ring_gemv(ring, P.ada_n, C, sA, warp, lane, [&](int n, float v) { P.fh_ada[n] = v + bias_at(P.ada_b, n); });
    gemv_direct(P.ip_w, C, ldim, sCur, gw, nw, lane, [&](int n, float v) { P.fh_x[n] = v + bias_at(P.ip_b, n); });
    for (int b = 0; b < P.depth; b++) {
      const FusedRes &rb = P.rb[b];
      const float *shift = P.fh_ada + (long)b * 3 * C, *scale = shift + C, *gate = shift + 2 * C;
      // ---- h2 = silu(mlp0(modulate(LN(fx)))) ----
      PHASE_BEGIN_LN(rb.lnw, rb.lnb, C);
      stage_ln(sA, red, P.fh_x, C, rb.lnw, rb.lnb, rb.eps, scale, shift);
      RSTAMP();
      ring_gemv(ring, P.hid, C, sA, warp, lane,
                [&](int n, float v) { P.fh_h2[n] = act_apply(v + bias_at(rb.mlp0b, n), ACT_SILU); });
      // ---- fx += gate * mlp2(h2) ----
      PHASE_BEGIN();
      stage_plain(sA, P.fh_h2, P.hid);
      RSTAMP();
      ring_gemv(ring, C, P.hid, sA, warp, lane, [&](int n, float v) {
        P.fh_x[n] = (v + bias_at(rb.mlp2b, n)) * ldcg1(gate + n) + ldcg1(P.fh_x + n);
      });
    }
    // ---- cur += final(modulate(LN(fx))) / steps ----
    const float *fshift = P.fh_ada + (long)P.depth * 3 * C, *fscale = fshift + C;
    PHASE_BEGIN();
    stage_ln(sA, red, P.fh_x, C, nullptr, nullptr, 1e-6f, fscale, fshift);
    RSTAMP();
    const bool last_it = it == lsd_steps - 1;
    ring_gemv(ring, ldim, C, sA, warp, lane, [&](int n, float v) {
      const float c = (v + bias_at(P.fin_b, n)) * inv_steps + ldcg1(P.fh_cur + n);
      P.fh_cur[n] = c;
      if (last_it)
        P.hist_w[off_b[0] + n] = c; // the new latent frame of this sequence
    });
  }
  RSTAMP();
#undef PHASE_BEGIN
#undef PHASE_BEGIN_LN
}

// CTAs (one per SM) and ring slots for an activation area of `act_bytes`; 0 when it cannot run.
int step_ring_grid(size_t act_bytes, int *slots, size_t *smem) {
  int dev = 0, sms = 0, optin = 0, occ = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  cudaFuncAttributes fa{};
  if (cudaFuncGetAttributes(&fa, (const void *)step_ring_kernel) != cudaSuccess)
    return 0;
  const long room = (long)optin - (long)fa.sharedSizeBytes - (long)act_bytes - 256;
  int S = (int)(room / RING_SLOT_BYTES);
  if (S > RING_MAX_SLOTS)
    S = RING_MAX_SLOTS;
  if (S < 8)
    return 0;
  const size_t total = (size_t)S * RING_SLOT_BYTES + act_bytes;
  if (cudaFuncSetAttribute((const void *)step_ring_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)total) !=
      cudaSuccess)
    return 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, (const void *)step_ring_kernel, RING_BLOCK, total) !=
          cudaSuccess ||
      occ < 1)
    return 0;
  *slots = S;
  *smem = total;
  return sms;
}

int step_ring_launch(const FusedParams *dp, int grid, size_t smem, int slots, const int32_t *row_slot,
                     const int32_t *row_pos, const int64_t *off_a, const int64_t *off_b, const int64_t *off_c,
                     int gather_frames, int gather_noise, const float *tvec, int lsd_steps, unsigned *barrier,
                     cudaStream_t st) {
  if (cudaMemsetAsync(barrier, 0, STEP_FUSED_BARRIER_BYTES, st) != cudaSuccess)
    return -1;
  void *args[] = {(void *)&dp,    (void *)&row_slot,      (void *)&row_pos,      (void *)&off_a, (void *)&off_b,
                  (void *)&off_c, (void *)&gather_frames, (void *)&gather_noise, (void *)&tvec,  (void *)&lsd_steps,
                  (void *)&slots};
  if (cudaLaunchCooperativeKernel((void *)step_ring_kernel, dim3(grid), dim3(RING_BLOCK), args, smem, st) != cudaSuccess)
    return -1;
  return 2; // memset + kernel
}

} // namespace ptts
