// step_fused.cu — one AR frame of ONE sequence (the batch-1 latency path) as a single
// persistent cooperative kernel: flow_lm.go:238-299 (NaN->BOS, input_linear, 6 transformer
// layers with RoPE + paged-KV append + attention, out_norm, EOS logit) and the LSD / flow
// head of flow_lm.go:311-353 + flow_net.go:314-356, 46 phases separated by grid barriers
// instead of 79 kernel launches.
//
// The frame is memory-bound (341 MB of fp32 weights, read once); what the multi-kernel
// version loses is the dependent-launch chain (3-5 us per kernel).  Here
//   * every warp requests the weight row of its first column of the NEXT phase before it
//     arrives at the barrier, so the HBM latency and part of the transfer overlap the
//     barrier and the staging of the next phase's activation vector;
//   * each CTA stages the (LayerNorm'd / modulated / activated) input vector of a phase in
//     shared memory once, then its warps stream weight rows against it;
//   * activations cross phases through L2 (`ld.global.cg`: L1 is not coherent across SMs).
// Grid = 2 CTAs of 128 threads per SM, launched cooperatively (co-residency is what makes
// the barrier legal); a spin that exceeds ~1 s raises an error flag instead of hanging.
//
// Two kernels share the weight streaming and staging code and differ in how a phase hands its
// output vector to the next one: step_fused_kernel (grid barrier, default) and
// step_fused_ll_kernel (flag-in-data words polled by the consumers, PTTS_FUSED_LL=1); both can
// bulk-prefetch the weight rows of later phases into L2 (PTTS_FUSED_L2_AHEAD).  Measured on
// B200: 266 / 277 us per frame, look-ahead slower in both (profiles/fused_matrix_r01e.txt) —
// the frame is bound by bytes / ~3 TB/s + ~3 us per phase whichever exchange is used; see
// DESIGN.md section 4 for the analysis and the planned TMA-fed shared-memory weight ring.
#include "step_fused.cuh"
#include "attend.cuh"
#include "gemm_epi.cuh"

namespace ptts {

namespace {

constexpr int FT = 128; // threads per CTA (= ATT_THREADS of attend_row)

__device__ __forceinline__ float4 ldg_w(const float *p) { // weights: read once, keep out of L1
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ float ldcg1(const float *p) { return __ldcg(p); }
__device__ __forceinline__ unsigned long long gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define FSTAMP()                                                               \
  do {                                                                         \
    if (P.dbg && blockIdx.x == 0 && threadIdx.x == 0 && nstamp < 500)          \
      P.dbg[nstamp++] = gtime();                                               \
  } while (0)
// Pull constant data (LayerNorm weights, KV rows of earlier tokens) into this SM's L1 ahead of
// the phase that reads it; costs no registers.
__device__ __forceinline__ void l1_prefetch(const void *p, size_t bytes) {
  if (!p)
    return;
  const char *c = reinterpret_cast<const char *>(p);
  for (size_t off = (size_t)threadIdx.x * 128; off < bytes; off += (size_t)FT * 128)
    asm volatile("prefetch.global.L1 [%0];" ::"l"(c + off));
}
__device__ __forceinline__ float bias_at(const float *b, int n) { return b ? __ldg(b + n) : 0.0f; }

// ---- L2 look-ahead ----
// The weights do not depend on the activations, so their HBM stream need not stop at the phase
// boundaries: while phase i computes, the rows this warp will read in phase i + l2_ahead are
// requested into L2 with bulk prefetches (one instruction per row, no registers, not in the
// L1tex FIFO).  The register loads of a phase then hit L2, and HBM stays busy across the
// exchanges between phases.  Weight phases in order: per layer in_proj, out_proj, linear1,
// linear2; cond_embed; per LSD step adaLN_all, per block mlp0, mlp2, final.
__device__ __forceinline__ bool phase_weights(const FusedParams &P, int idx, int lsd_steps, const float *&W, int &N,
                                              int &K) {
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
    W = P.ada_w; N = P.ada_n; K = P.C;
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
__device__ __forceinline__ void l2_prefetch_phase(const FusedParams &P, int idx, int lsd_steps, int gw, int nw, int lane) {
  const float *W;
  int N, K;
  if (!phase_weights(P, idx, lsd_steps, W, N, K))
    return;
  for (int n = gw + lane * nw; n < N; n += 32 * nw) // lane i: the warp's i-th column, the whole row
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(W + (long)n * K), "r"(K * 4) : "memory");
}

// Grid barrier on 16 striped counters (one 128-byte line each).  Same-address atomics
// serialise in L2 (~27 cycles each): one counter costs ~4 us for 296 CTAs, a two-level tree
// two dependent round trips, per-CTA flags polled by everybody flood L2.  Here CTA b adds to
// counter b % 16 with a no-return reduction (~18 arrivals per counter, in parallel) and
// lanes 0..15 of warp 0 poll one counter each.  Counters are monotonic within a launch (the
// host zeroes them before it).
constexpr int GB_NCTR = 16;
constexpr int GB_STRIDE = 32; // words between counters
struct GridBar {
  unsigned *ctr;
  int *err;
  unsigned phase, mine; // arrivals per phase on the counter this lane polls
  bool dead;
};
__device__ __forceinline__ GridBar make_bar(unsigned *ctr, int *err) {
  GridBar b;
  b.ctr = ctr;
  b.err = err;
  b.phase = 0;
  b.dead = false;
  const unsigned c = threadIdx.x & (GB_NCTR - 1);
  b.mine = (gridDim.x + GB_NCTR - 1 - c) / GB_NCTR; // CTAs with blockIdx % 16 == c
  return b;
}
// Arrive: publish this CTA's writes and count it in.  Must come BEFORE the weight prefetch of
// the next phase: the gpu-scope fence waits for the issuing thread's outstanding loads.
__device__ __forceinline__ void grid_arrive(GridBar &b) {
  b.phase++;
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    asm volatile("red.relaxed.gpu.global.add.u32 [%0], 1;" ::"l"(b.ctr + GB_STRIDE * (blockIdx.x & (GB_NCTR - 1)))
                 : "memory");
  }
}
// Wait: acquire loads only (no fence: the prefetch loads stay in flight); the CTA barrier
// joins the pollers' acquires.
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
        atomicExch(b.err, 1); // somebody is missing: give up instead of hanging the GPU
        break;
      }
    }
  }
  __syncthreads();
  if (!b.dead && *reinterpret_cast<volatile int *>(b.err) != 0)
    b.dead = true;
}

// ---- weight streaming ----
// A warp's share of a phase is a list of segments: (column n, 1024-wide k chunk c), columns
// n = gw, gw + nw, ...  Segments are processed in batches of 4 (32 float4 = 8 KB per warp
// in flight); the first two segments are requested before the grid barrier, the rest of
// the first batch right after it, so they land while the CTA stages its activation vector.
constexpr int SEG_K = 1024;
struct Pref {
  float4 w[2][8];
};
__device__ __forceinline__ void seg_load(float4 (&w)[8], const float *W, int N, int K, int kc, int gw, int nw, int seg,
                                         int lane) {
  const int n = gw + (seg / kc) * nw;
  const int k0 = (seg % kc) * SEG_K;
#pragma unroll
  for (int u = 0; u < 8; u++) {
    const int k = k0 + lane * 4 + 128 * u;
    w[u] = (n < N && k < K) ? ldg_w(W + (long)n * K + k) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
}
__device__ __forceinline__ void prefetch_w(Pref &p, const float *W, int N, int K, int gw, int nw, int lane) {
  const int kc = (K + SEG_K - 1) / SEG_K;
  seg_load(p.w[0], W, N, K, kc, gw, nw, 0, lane);
  seg_load(p.w[1], W, N, K, kc, gw, nw, 1, lane);
}
__device__ __forceinline__ float dot4(const float4 &a, const float4 &w, float acc) {
  acc = fmaf(a.x, w.x, acc);
  acc = fmaf(a.y, w.y, acc);
  acc = fmaf(a.z, w.z, acc);
  acc = fmaf(a.w, w.w, acc);
  return acc;
}
__device__ __forceinline__ float seg_dot(const float4 (&w)[8], const float *sA, int K, int k0, int lane) {
  float acc = 0.0f;
#pragma unroll
  for (int u = 0; u < 8; u++) {
    const int k = k0 + lane * 4 + 128 * u;
    if (k < K)
      acc = dot4(*reinterpret_cast<const float4 *>(sA + k), w[u], acc);
  }
  return acc;
}
// Second half of the first batch: called right after the barrier, before the staging.
struct Pref2 {
  float4 w[2][8];
};
__device__ __forceinline__ void prefetch_w2(Pref2 &p, const float *W, int N, int K, int gw, int nw, int lane) {
  const int kc = (K + SEG_K - 1) / SEG_K;
  seg_load(p.w[0], W, N, K, kc, gw, nw, 2, lane);
  seg_load(p.w[1], W, N, K, kc, gw, nw, 3, lane);
}
// y[n] = W[n,:K] . a for this warp's columns; a is in shared memory; epi(n, y) runs on lane 0.
template <class Epi>
__device__ __forceinline__ void gemv_cols(const Pref &p, const Pref2 &p2, const float *W, int N, int K, const float *sA,
                                          int gw, int nw, int lane, Epi epi) {
  const int kc = (K + SEG_K - 1) / SEG_K;
  const int ncols = gw < N ? (N - gw + nw - 1) / nw : 0;
  const int nseg = ncols * kc;
  float colacc = 0.0f;
  auto finish = [&](int seg, float part) { // partial of segment `seg`; emit the column at its last chunk
    colacc += part;
    if ((seg % kc) == kc - 1) {
      float v = colacc;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1)
        v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == 0)
        epi(gw + (seg / kc) * nw, v);
      colacc = 0.0f;
    }
  };
  // first batch from the prefetched registers
  if (nseg > 0) finish(0, seg_dot(p.w[0], sA, K, 0, lane));
  if (nseg > 1) finish(1, seg_dot(p.w[1], sA, K, (1 % kc) * SEG_K, lane));
  if (nseg > 2) finish(2, seg_dot(p2.w[0], sA, K, (2 % kc) * SEG_K, lane));
  if (nseg > 3) finish(3, seg_dot(p2.w[1], sA, K, (3 % kc) * SEG_K, lane));
  for (int s0 = 4; s0 < nseg; s0 += 4) {
    float4 w[4][8];
#pragma unroll
    for (int j = 0; j < 4; j++)
      seg_load(w[j], W, N, K, kc, gw, nw, s0 + j, lane); // beyond nseg: n >= N -> zeros, no loads
#pragma unroll
    for (int j = 0; j < 4; j++)
      if (s0 + j < nseg)
        finish(s0 + j, seg_dot(w[j], sA, K, ((s0 + j) % kc) * SEG_K, lane));
  }
}

// K == SEG_K / 2 (the adaLN projection, 10240 x 512): a 1024-wide register segment would be
// half empty and a warp's eight or nine columns would take three dependent batches.  Two
// consecutive rows are contiguous in memory, so the matrix is read as [N/2, 2K]: one segment =
// one row pair, the activation vector is staged twice back to back (sA2[k] == sA2[k + K]) and the
// two halves of the lane's partial sums go to the two rows.  Per-lane summation order and the
// shuffles are those of gemv_cols, so the results are bit-identical.
template <class Epi>
__device__ __forceinline__ void gemv_pairs(const Pref &p, const Pref2 &p2, const float *W, int NP, const float *sA2,
                                           int gw, int nw, int lane, Epi epi) {
  const int ncols = gw < NP ? (NP - gw + nw - 1) / nw : 0; // row pairs of this warp
  auto finish = [&](int seg, const float4 (&w)[8]) {
    float lo = 0.0f, hi = 0.0f;
#pragma unroll
    for (int u = 0; u < 4; u++)
      lo = dot4(*reinterpret_cast<const float4 *>(sA2 + lane * 4 + 128 * u), w[u], lo);
#pragma unroll
    for (int u = 4; u < 8; u++)
      hi = dot4(*reinterpret_cast<const float4 *>(sA2 + lane * 4 + 128 * u), w[u], hi);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      lo += __shfl_xor_sync(0xffffffffu, lo, o);
      hi += __shfl_xor_sync(0xffffffffu, hi, o);
    }
    if (lane == 0) {
      const int m = gw + seg * nw;
      epi(2 * m, lo);
      epi(2 * m + 1, hi);
    }
  };
  if (ncols > 0) finish(0, p.w[0]);
  if (ncols > 1) finish(1, p.w[1]);
  if (ncols > 2) finish(2, p2.w[0]);
  if (ncols > 3) finish(3, p2.w[1]);
  for (int s0 = 4; s0 < ncols; s0 += 4) {
    float4 w[4][8];
#pragma unroll
    for (int j = 0; j < 4; j++)
      seg_load(w[j], W, NP, SEG_K, 1, gw, nw, s0 + j, lane); // beyond ncols: n >= NP -> zeros, no loads
#pragma unroll
    for (int j = 0; j < 4; j++)
      if (s0 + j < ncols)
        finish(s0 + j, w[j]);
  }
}

__device__ __forceinline__ float block_sum(float v, float *red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
    v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0)
    red[threadIdx.x >> 5] = v;
  __syncthreads();
  return (red[0] + red[1]) + (red[2] + red[3]);
}
// sA[0:K] = src (through L2)
__device__ __forceinline__ void stage_plain(float *sA, const float *src, int K) {
  for (int i = threadIdx.x; i < K; i += FT)
    sA[i] = ldcg1(src + i);
  __syncthreads();
}
// sA = (LN(src) * w + b) [* (1 + scale) + shift]  (linear.go:265-329, tensor_util.go:175-193)
__device__ __forceinline__ void stage_ln(float *sA, float *red, const float *src, int K, const float *w, const float *b,
                                         float eps, const float *scale, const float *shift) {
  float s = 0.0f;
  for (int i = threadIdx.x; i < K; i += FT) {
    const float v = ldcg1(src + i);
    sA[i] = v;
    s += v;
  }
  const float mean = block_sum(s, red) / (float)K;
  float q = 0.0f;
  for (int i = threadIdx.x; i < K; i += FT) {
    const float d = sA[i] - mean;
    q += d * d;
  }
  const float inv = 1.0f / sqrtf(block_sum(q, red) / (float)K + eps);
  for (int i = threadIdx.x; i < K; i += FT) {
    float t = (sA[i] - mean) * inv;
    if (w)
      t *= __ldg(w + i);
    if (b)
      t += __ldg(b + i);
    if (scale)
      t = t * (ldcg1(scale + i) + 1.0f) + ldcg1(shift + i);
    sA[i] = t;
  }
  __syncthreads();
}

} // namespace

__global__ void __launch_bounds__(FT, 2)
    step_fused_kernel(const FusedParams *__restrict__ Pp, const int32_t *row_slot, const int32_t *row_pos,
                      const int64_t *off_a, const int64_t *off_b, const int64_t *off_c, int gather_frames,
                      int gather_noise, const float *tvec, int lsd_steps) {
  extern __shared__ float dsm[]; // [max(F, max_keys)] activation vector / attention scores
  __shared__ float red[FT / 32];
  __shared__ float4 pv[8 * 16];
  __shared__ __align__(16) float q_s[HD];
  const FusedParams &P = *Pp;
  const int lane = threadIdx.x & 31;
  const int gw = (blockIdx.x * FT + threadIdx.x) >> 5;
  const int nw = (gridDim.x * FT) >> 5;
  GridBar bar = make_bar(P.barrier, P.error);
  float *sA = dsm;
  const int D = P.D, F = P.F, C = P.C, ldim = P.ldim;
  const int slot = row_slot[0], pos = row_pos[0];
  Pref pf;
  Pref2 pf2;
  int nstamp = 0;
  FSTAMP();
  int wph = 0; // index of the weight phase being started (see phase_weights)
#define L2_AHEAD()                                                             \
  do {                                                                         \
    if (P.l2_ahead > 0)                                                        \
      l2_prefetch_phase(P, wph + P.l2_ahead, lsd_steps, gw, nw, lane);         \
    wph++;                                                                     \
  } while (0)
  for (int a = 0; a < P.l2_ahead; a++)
    l2_prefetch_phase(P, a, lsd_steps, gw, nw, lane);
// publish the previous phase, request the first weight segments of the next one, wait for
// everybody, request the rest of the first batch (they land while the vector is staged)
#define PHASE_BEGIN(W, N, K) PHASE_BEGIN_LN(W, N, K, nullptr, nullptr, 0)
#define PHASE_BEGIN_LN(W, N, K, LW, LB, KLN)                                   \
  do {                                                                         \
    grid_arrive(bar);                                                          \
    prefetch_w(pf, W, N, K, gw, nw, lane);                                     \
    L2_AHEAD();                                                                \
    l1_prefetch(LW, (size_t)(KLN) * 4);                                        \
    l1_prefetch(LB, (size_t)(KLN) * 4);                                        \
    grid_wait(bar);                                                            \
    prefetch_w2(pf2, W, N, K, gw, nw, lane);                                   \
    FSTAMP();                                                                  \
  } while (0)
  auto gemv_now = [&](const float *W, int N, int K, const float *a, int gw2, auto epi) {
    Pref q;
    Pref2 q2;
    prefetch_w(q, W, N, K, gw2, nw, lane);
    prefetch_w2(q2, W, N, K, gw2, nw, lane);
    gemv_cols(q, q2, W, N, K, a, gw2, nw, lane, epi);
  };

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
    if (blockIdx.x == 0 && threadIdx.x < ldim) { // the flow head starts from the noise
      const float *nz = gather_noise ? P.noise_pool + off_c[0] : P.fh_noise;
      P.fh_cur[threadIdx.x] = ldcg1(nz + threadIdx.x);
    }
    __syncthreads();
    gemv_now(P.in_w, D, ldim, sA, gw, [&](int n, float v) { P.x[n] = v + bias_at(P.in_b, n); });
  }
  for (int l = 0; l < P.L; l++) {
    const FusedLayer &ly = P.layer[l];
    // ---- qkv = in_proj(LN1(x)) ----
    PHASE_BEGIN_LN(ly.in_proj, 3 * D, D, ly.n1w, ly.n1b, D);
    stage_ln(sA, red, P.x, D, ly.n1w, ly.n1b, ly.eps1, nullptr, nullptr);
    FSTAMP();
    gemv_cols(pf, pf2, ly.in_proj, 3 * D, D, sA, gw, nw, lane, [&](int n, float v) { P.qkv[n] = v; });
    // ---- RoPE, KV append, attention: one CTA per head; the others already hold the first
    //      out_proj segments when they reach the next barrier ----
    grid_arrive(bar);
    prefetch_w(pf, ly.out_proj, D, D, gw, nw, lane);
    L2_AHEAD();
    grid_wait(bar);
    FSTAMP();
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
      __syncthreads();
      const __half *pool = static_cast<const __half *>(kv.pool);
      auto keyp = [&](int j) { return pool + (long)pt[j / PT] * page_stride + koff + (long)(j % PT) * HD; };
      auto valp = [&](int j) { return pool + (long)pt[j / PT] * page_stride + voff + (long)(j % PT) * HD; };
      if (PT == 16) // one pass, one block barrier (attend.cuh): this phase is a latency chain on 16 CTAs
        attend_row_flash16(q_s, pos + 1, pool + koff, pool + voff, pt, page_stride, P.att + h * HD, dsm, SplitOut(), 0, 0);
      else
        attend_row(q_s, pos + 1, keyp, valp, P.att + h * HD, dsm, red, pv, SplitOut(), 0, 0);
    }
    // ---- x += out_proj(att) ----
    grid_arrive(bar);
    grid_wait(bar);
    prefetch_w2(pf2, ly.out_proj, D, D, gw, nw, lane);
    FSTAMP();
    stage_plain(sA, P.att, D);
    FSTAMP();
    gemv_cols(pf, pf2, ly.out_proj, D, D, sA, gw, nw, lane, [&](int n, float v) { P.x[n] = ldcg1(P.x + n) + v; });
    // ---- ff = gelu(linear1(LN2(x))) ----
    PHASE_BEGIN_LN(ly.l1, F, D, ly.n2w, ly.n2b, D);
    stage_ln(sA, red, P.x, D, ly.n2w, ly.n2b, ly.eps2, nullptr, nullptr);
    FSTAMP();
    gemv_cols(pf, pf2, ly.l1, F, D, sA, gw, nw, lane, [&](int n, float v) { P.ff[n] = act_apply(v, ACT_GELU); });
    // ---- x += linear2(ff) ----
    PHASE_BEGIN(ly.l2, D, F);
    stage_plain(sA, P.ff, F);
    FSTAMP();
    gemv_cols(pf, pf2, ly.l2, D, F, sA, gw, nw, lane, [&](int n, float v) { P.x[n] = ldcg1(P.x + n) + v; });
  }
  // ---- last = out_norm(x); eos = last . w_eos + b; cproj = cond_embed(last) ----
  PHASE_BEGIN_LN(P.ce_w, C, D, P.on_w, P.on_b, D);
  stage_ln(sA, red, P.x, D, P.on_w, P.on_b, P.on_eps, nullptr, nullptr);
  FSTAMP();
  if (blockIdx.x == 0)
    for (int i = threadIdx.x; i < D; i += FT)
      P.fh_last[i] = sA[i];
  gemv_cols(pf, pf2, P.ce_w, C, D, sA, gw, nw, lane, [&](int n, float v) { P.fh_cproj[n] = v + bias_at(P.ce_b, n); });
  if (gw == nw - 1) // the EOS logit is one more column, on the last warp
    gemv_now(P.eos_w, 1, D, sA, 0, [&](int, float v) { P.fh_eos[0] = v + P.eos_b; });
  const float inv_steps = 1.0f / (float)lsd_steps;
  for (int it = 0; it < lsd_steps; it++) {
    // ---- ada = adaLN_all(silu(cproj + t_emb)); fx = input_proj(cur) ----
    const bool pair_ada = 2 * C == SEG_K && (P.ada_n & 1) == 0; // see gemv_pairs
    float *sCur = sA + (pair_ada ? 2 * C : C);
    if (pair_ada)
      PHASE_BEGIN(P.ada_w, P.ada_n / 2, 2 * C);
    else
      PHASE_BEGIN(P.ada_w, P.ada_n, C);
    for (int i = threadIdx.x; i < C; i += FT) {
      float y = ldcg1(P.fh_cproj + i) + __ldg(tvec + (long)it * C + i);
      y = y / (1.0f + expf(-y));
      sA[i] = y;
      if (pair_ada)
        sA[C + i] = y;
    }
    for (int i = threadIdx.x; i < ldim; i += FT)
      sCur[i] = ldcg1(P.fh_cur + i);
    __syncthreads();
    FSTAMP();
    if (pair_ada)
      gemv_pairs(pf, pf2, P.ada_w, P.ada_n / 2, sA, gw, nw, lane,
                 [&](int n, float v) { P.fh_ada[n] = v + bias_at(P.ada_b, n); });
    else
      gemv_cols(pf, pf2, P.ada_w, P.ada_n, C, sA, gw, nw, lane,
                [&](int n, float v) { P.fh_ada[n] = v + bias_at(P.ada_b, n); });
    gemv_now(P.ip_w, C, ldim, sCur, gw, [&](int n, float v) { P.fh_x[n] = v + bias_at(P.ip_b, n); });
    for (int b = 0; b < P.depth; b++) {
      const FusedRes &rb = P.rb[b];
      const float *shift = P.fh_ada + (long)b * 3 * C, *scale = shift + C, *gate = shift + 2 * C;
      // ---- h2 = silu(mlp0(modulate(LN(fx)))) ----
      PHASE_BEGIN_LN(rb.mlp0w, P.hid, C, rb.lnw, rb.lnb, C);
      stage_ln(sA, red, P.fh_x, C, rb.lnw, rb.lnb, rb.eps, scale, shift);
      FSTAMP();
      gemv_cols(pf, pf2, rb.mlp0w, P.hid, C, sA, gw, nw, lane,
                [&](int n, float v) { P.fh_h2[n] = act_apply(v + bias_at(rb.mlp0b, n), ACT_SILU); });
      // ---- fx += gate * mlp2(h2) ----
      PHASE_BEGIN(rb.mlp2w, C, P.hid);
      stage_plain(sA, P.fh_h2, P.hid);
      FSTAMP();
      gemv_cols(pf, pf2, rb.mlp2w, C, P.hid, sA, gw, nw, lane, [&](int n, float v) {
        P.fh_x[n] = (v + bias_at(rb.mlp2b, n)) * ldcg1(gate + n) + ldcg1(P.fh_x + n);
      });
    }
    // ---- cur += final(modulate(LN(fx))) / steps ----
    const float *fshift = P.fh_ada + (long)P.depth * 3 * C, *fscale = fshift + C;
    PHASE_BEGIN(P.fin_w, ldim, C);
    stage_ln(sA, red, P.fh_x, C, nullptr, nullptr, 1e-6f, fscale, fshift);
    FSTAMP();
    const bool last_it = it == lsd_steps - 1;
    gemv_cols(pf, pf2, P.fin_w, ldim, C, sA, gw, nw, lane, [&](int n, float v) {
      const float c = (v + bias_at(P.fin_b, n)) * inv_steps + ldcg1(P.fh_cur + n);
      P.fh_cur[n] = c;
      if (last_it)
        P.hist_w[off_b[0] + n] = c; // the new latent frame of this sequence
    });
  }
  FSTAMP();
#undef PHASE_BEGIN
#undef PHASE_BEGIN_LN
#undef L2_AHEAD
}

// ===========================================================================================
// Flag-in-data variant.  The grid barrier above costs three dependent L2 round trips per
// phase (fence + arrive, poll, staging loads).  Here every activation that crosses CTAs is an
// 8-byte (value, tag) word written with ONE relaxed 64-bit store by the lane that produced
// it; consumers poll the data itself until every word carries the tag of the producing phase
// (8-byte accesses are single-copy atomic, so a word is either old or new, never torn).  No
// counters, no fences, no arrive: producer store -> L2 -> consumer load is the whole
// hand-over, and a consumer only waits for the words it reads (an attention CTA for the 192
// q/k/v columns of its head).  Tags are `epoch + phase index`, unique per phase and launch;
// `epoch` lives in device memory so the launch can be replayed from a CUDA graph.
//
// Why one buffer per tensor is enough (no double buffering).  Writes of phase p happen after
// the writer staged the outputs of phase p-1 it needs; by induction over the phases every
// CTA that read anything in a phase q < p has finished those reads (a CTA only reads in a
// phase in which it owns columns, and it publishes its columns after its reads).  What a
// phase must not do is write a tensor that the SAME phase stages; the phase list below never
// does (x is updated in place by out_proj / linear2, which stage att / ff; fx by mlp2, which
// stages h2; cur by the final layer, which stages fx and ada).
namespace {

struct Poll {
  int *err;
  int sleep_ns;
  bool dead;
};
__device__ __forceinline__ void ll_st(uint2 *buf, int n, float v, unsigned tag) {
  asm volatile("st.relaxed.gpu.global.v2.u32 [%0], {%1, %2};" ::"l"(buf + n), "r"(__float_as_uint(v)), "r"(tag)
               : "memory");
}
__device__ __forceinline__ uint4 ll_ld2(const uint2 *p) { // two words; each half is atomic on its own
  uint4 r;
  asm volatile("ld.relaxed.gpu.global.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p)
               : "memory");
  return r;
}
__device__ __forceinline__ uint2 ll_ld1(const uint2 *p) {
  uint2 r;
  asm volatile("ld.relaxed.gpu.global.v2.u32 {%0, %1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p) : "memory");
  return r;
}
// After a failed poll: back off, give up after ~seconds (a missing producer must not hang the GPU).
__device__ __forceinline__ bool poll_fail(Poll &pl, int &spins) {
  if (pl.sleep_ns)
    __nanosleep(pl.sleep_ns);
  if (++spins > 3000000 || ((spins & 1023) == 0 && *reinterpret_cast<volatile int *>(pl.err) != 0)) {
    atomicExch(pl.err, 1);
    pl.dead = true;
    return true;
  }
  return false;
}
__device__ __forceinline__ float ll_get1(const uint2 *buf, int n, unsigned tag, Poll &pl) {
  uint2 r = ll_ld1(buf + n);
  int spins = 0;
  while (r.y != tag && !pl.dead) {
    if (poll_fail(pl, spins))
      break;
    r = ll_ld1(buf + n);
  }
  return __uint_as_float(r.x);
}
__device__ __forceinline__ float2 ll_get2(const uint2 *buf, int n, unsigned tag, Poll &pl) {
  uint4 r = ll_ld2(buf + n);
  int spins = 0;
  while ((r.y != tag || r.w != tag) && !pl.dead) {
    if (poll_fail(pl, spins))
      break;
    r = ll_ld2(buf + n);
  }
  return make_float2(__uint_as_float(r.x), __uint_as_float(r.z));
}
// sA[0:K] <- values of buf[0:K] once they all carry `tag` (K even).  Thread t owns the pairs
// 2t + 2*FT*u; up to four 16-byte loads are in flight per thread before the first check.
__device__ __forceinline__ void ll_stage_raw(float *sA, const uint2 *buf, int K, unsigned tag, Poll &pl) {
  for (int i0 = 2 * threadIdx.x; i0 < K; i0 += 8 * FT) {
    uint4 r[4];
#pragma unroll
    for (int u = 0; u < 4; u++) {
      const int i = i0 + 2 * FT * u;
      if (i < K)
        r[u] = ll_ld2(buf + i);
    }
#pragma unroll
    for (int u = 0; u < 4; u++) {
      const int i = i0 + 2 * FT * u;
      if (i < K) {
        int spins = 0;
        while ((r[u].y != tag || r[u].w != tag) && !pl.dead) {
          if (poll_fail(pl, spins))
            break;
          r[u] = ll_ld2(buf + i);
        }
        *reinterpret_cast<float2 *>(sA + i) = make_float2(__uint_as_float(r[u].x), __uint_as_float(r[u].z));
      }
    }
  }
}
__device__ __forceinline__ void ll_stage_plain(float *sA, const uint2 *buf, int K, unsigned tag, Poll &pl) {
  ll_stage_raw(sA, buf, K, tag, pl);
  __syncthreads();
}
// As stage_ln, from (value, tag) words; the adaLN scale / shift vectors (also tagged) are staged
// behind the row at sA[K:2K] / sA[2K:3K].  Every pass touches only the thread's own pairs.
__device__ __forceinline__ void ll_stage_ln(float *sA, float *red, const uint2 *src, int K, unsigned tag, const float *w,
                                            const float *b, float eps, const uint2 *scale, const uint2 *shift,
                                            unsigned ada_tag, Poll &pl) {
  ll_stage_raw(sA, src, K, tag, pl);
  if (scale) {
    ll_stage_raw(sA + K, scale, K, ada_tag, pl);
    ll_stage_raw(sA + 2 * K, shift, K, ada_tag, pl);
  }
  float s = 0.0f;
  for (int i = 2 * threadIdx.x; i < K; i += 2 * FT)
    s += sA[i] + sA[i + 1];
  const float mean = block_sum(s, red) / (float)K;
  float q = 0.0f;
  for (int i = 2 * threadIdx.x; i < K; i += 2 * FT) {
    const float d0 = sA[i] - mean, d1 = sA[i + 1] - mean;
    q += d0 * d0 + d1 * d1;
  }
  const float inv = 1.0f / sqrtf(block_sum(q, red) / (float)K + eps);
  for (int i0 = 2 * threadIdx.x; i0 < K; i0 += 2 * FT) {
#pragma unroll
    for (int e = 0; e < 2; e++) {
      const int i = i0 + e;
      float t = (sA[i] - mean) * inv;
      if (w)
        t *= __ldg(w + i);
      if (b)
        t += __ldg(b + i);
      if (scale)
        t = t * (sA[K + i] + 1.0f) + sA[2 * K + i];
      sA[i] = t;
    }
  }
  __syncthreads();
}

} // namespace

__global__ void __launch_bounds__(FT, 2)
    step_fused_ll_kernel(const FusedParams *__restrict__ Pp, const int32_t *row_slot, const int32_t *row_pos,
                         const int64_t *off_a, const int64_t *off_b, const int64_t *off_c, int gather_frames,
                         int gather_noise, const float *tvec, int lsd_steps) {
  extern __shared__ float dsm[]; // activation vector (+ adaLN scale / shift) / attention scores
  __shared__ float red[FT / 32];
  __shared__ float4 pv[8 * 16];
  __shared__ __align__(16) float q_s[HD];
  const FusedParams &P = *Pp;
  const int lane = threadIdx.x & 31;
  const int gw = (blockIdx.x * FT + threadIdx.x) >> 5;
  const int gw0 = (blockIdx.x * FT) >> 5; // first warp of this CTA
  const int nw = (gridDim.x * FT) >> 5;
  float *sA = dsm;
  const int D = P.D, F = P.F, C = P.C, ldim = P.ldim;
  const int slot = row_slot[0], pos = row_pos[0];
  Poll pl{P.error, P.poll_ns, false};
  const unsigned base = __ldcg(P.epoch);
  unsigned T = base; // tag of the phase being produced
  Pref pf;
  Pref2 pf2;
  int nstamp = 0;
  FSTAMP();
  auto gemv_now = [&](const float *W, int N, int K, const float *a, int gw2, auto epi) {
    Pref q;
    Pref2 q2;
    prefetch_w(q, W, N, K, gw2, nw, lane);
    prefetch_w2(q2, W, N, K, gw2, nw, lane);
    gemv_cols(q, q2, W, N, K, a, gw2, nw, lane, epi);
  };
  // the first weight batch of a phase is requested before anything is polled: it is in
  // flight while the producers of the previous phase finish
#define L2_AHEAD()                                                             \
  do {                                                                         \
    if (P.l2_ahead > 0)                                                        \
      l2_prefetch_phase(P, wph + P.l2_ahead, lsd_steps, gw, nw, lane);         \
    wph++;                                                                     \
  } while (0)
#define LL_PREFETCH(W, N, K, LW, LB, KLN)                                      \
  do {                                                                         \
    T++;                                                                       \
    prefetch_w(pf, W, N, K, gw, nw, lane);                                     \
    prefetch_w2(pf2, W, N, K, gw, nw, lane);                                   \
    L2_AHEAD();                                                                \
    l1_prefetch(LW, (size_t)(KLN) * 4);                                        \
    l1_prefetch(LB, (size_t)(KLN) * 4);                                        \
    __syncthreads(); /* the previous phase's dots are done with sA */          \
  } while (0)
  const float *noise0 = gather_noise ? P.noise_pool + off_c[0] : P.fh_noise;
  int wph = 0; // index of the weight phase being started (see phase_weights)
  for (int a = 0; a < P.l2_ahead; a++)
    l2_prefetch_phase(P, a, lsd_steps, gw, nw, lane);

  // ---- phase 0: x = input_linear(nan_to_bos(frame)) ----
  if (gw0 < D || blockIdx.x == 0) {
    const float *frame = gather_frames ? P.hist + off_a[0] : P.fh_frames;
    if (threadIdx.x < ldim) {
      const float v = ldcg1(frame + threadIdx.x);
      const float u = isnan(v) ? __ldg(P.bos + threadIdx.x) : v;
      sA[threadIdx.x] = u;
      if (blockIdx.x == 0)
        P.fh_in[threadIdx.x] = u;
    }
    __syncthreads();
    gemv_now(P.in_w, D, ldim, sA, gw, [&](int n, float v) { ll_st(P.ll_x, n, v + bias_at(P.in_b, n), T); });
  }
  unsigned t_x = T;
  for (int l = 0; l < P.L; l++) {
    const FusedLayer &ly = P.layer[l];
    // ---- qkv = in_proj(LN1(x)) ----
    LL_PREFETCH(ly.in_proj, 3 * D, D, ly.n1w, ly.n1b, D);
    if (gw0 < 3 * D) {
      ll_stage_ln(sA, red, P.ll_x, D, t_x, ly.n1w, ly.n1b, ly.eps1, nullptr, nullptr, 0, pl);
      FSTAMP();
      gemv_cols(pf, pf2, ly.in_proj, 3 * D, D, sA, gw, nw, lane, [&](int n, float v) { ll_st(P.ll_qkv, n, v, T); });
    }
    const unsigned t_qkv = T;
    // ---- RoPE, KV append, attention: one CTA per head, which waits only for the 192 q/k/v
    //      words of its head; everybody else goes straight on to poll att ----
    T++;
    prefetch_w(pf, ly.out_proj, D, D, gw, nw, lane);
    L2_AHEAD();
    __syncthreads(); // the in_proj dots of this CTA are done with sA (= the scores below)
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
        const float2 qq = ll_get2(P.ll_qkv, h * HD + 2 * j, t_qkv, pl);
        const float2 kk = ll_get2(P.ll_qkv, D + h * HD + 2 * j, t_qkv, pl);
        q_s[2 * j] = rope_even(qq.x, qq.y, c, s);
        q_s[2 * j + 1] = rope_odd(qq.x, qq.y, c, s);
        __half *dst = page + koff + (long)tok * HD + 2 * j;
        kv_store(dst, rope_even(kk.x, kk.y, c, s));
        kv_store(dst + 1, rope_odd(kk.x, kk.y, c, s));
      } else if (threadIdx.x < half + HD) {
        const int d = threadIdx.x - half;
        kv_store(page + voff + (long)tok * HD + d, ll_get1(P.ll_qkv, 2 * D + h * HD + d, t_qkv, pl));
      }
      __syncthreads();
      FSTAMP();
      const __half *pool = static_cast<const __half *>(kv.pool);
      auto keyp = [&](int j) { return pool + (long)pt[j / PT] * page_stride + koff + (long)(j % PT) * HD; };
      auto valp = [&](int j) { return pool + (long)pt[j / PT] * page_stride + voff + (long)(j % PT) * HD; };
      if (PT == 16)
        attend_row_flash16(q_s, pos + 1, pool + koff, pool + voff, pt, page_stride, (float *)nullptr, dsm, SplitOut(), 0,
                           0, P.ll_att + h * HD, T);
      else
        attend_row(q_s, pos + 1, keyp, valp, nullptr, dsm, red, pv, SplitOut(), 0, 0, P.ll_att + h * HD, T);
      __syncthreads(); // scores (dsm) are restaged as the activation vector below
    }
    const unsigned t_att = T;
    // ---- x += out_proj(att) ----
    T++;
    prefetch_w2(pf2, ly.out_proj, D, D, gw, nw, lane);
    FSTAMP();
    if (gw0 < D) {
      ll_stage_plain(sA, P.ll_att, D, t_att, pl);
      FSTAMP();
      float pre = 0.0f; // the residual of this warp's first column, in flight during the dot
      if (lane == 0 && gw < D)
        pre = ll_get1(P.ll_x, gw, t_x, pl);
      gemv_cols(pf, pf2, ly.out_proj, D, D, sA, gw, nw, lane, [&](int n, float v) {
        const float xo = n == gw ? pre : ll_get1(P.ll_x, n, t_x, pl);
        ll_st(P.ll_x, n, xo + v, T);
      });
    }
    t_x = T;
    // ---- ff = gelu(linear1(LN2(x))) ----
    LL_PREFETCH(ly.l1, F, D, ly.n2w, ly.n2b, D);
    if (gw0 < F) {
      ll_stage_ln(sA, red, P.ll_x, D, t_x, ly.n2w, ly.n2b, ly.eps2, nullptr, nullptr, 0, pl);
      FSTAMP();
      gemv_cols(pf, pf2, ly.l1, F, D, sA, gw, nw, lane,
                [&](int n, float v) { ll_st(P.ll_ff, n, act_apply(v, ACT_GELU), T); });
    }
    const unsigned t_ff = T;
    // ---- x += linear2(ff) ----
    LL_PREFETCH(ly.l2, D, F, nullptr, nullptr, 0);
    if (gw0 < D) {
      ll_stage_plain(sA, P.ll_ff, F, t_ff, pl);
      FSTAMP();
      float pre = 0.0f;
      if (lane == 0 && gw < D)
        pre = ll_get1(P.ll_x, gw, t_x, pl);
      gemv_cols(pf, pf2, ly.l2, D, F, sA, gw, nw, lane, [&](int n, float v) {
        const float xo = n == gw ? pre : ll_get1(P.ll_x, n, t_x, pl);
        ll_st(P.ll_x, n, xo + v, T);
      });
    }
    t_x = T;
  }
  // ---- last = out_norm(x); eos = last . w_eos + b; cproj = cond_embed(last) ----
  LL_PREFETCH(P.ce_w, C, D, P.on_w, P.on_b, D);
  if (gw0 < C) { // CTA 0 always owns column 0
    ll_stage_ln(sA, red, P.ll_x, D, t_x, P.on_w, P.on_b, P.on_eps, nullptr, nullptr, 0, pl);
    FSTAMP();
    if (blockIdx.x == 0)
      for (int i = threadIdx.x; i < D; i += FT)
        P.fh_last[i] = sA[i];
    gemv_cols(pf, pf2, P.ce_w, C, D, sA, gw, nw, lane,
              [&](int n, float v) { ll_st(P.ll_cproj, n, v + bias_at(P.ce_b, n), T); });
    if (gw == 0) // the EOS logit is one more column
      gemv_now(P.eos_w, 1, D, sA, 0, [&](int, float v) { P.fh_eos[0] = v + P.eos_b; });
  }
  const unsigned t_cproj = T;
  unsigned t_cur = 0;
  const float inv_steps = 1.0f / (float)lsd_steps;
  for (int it = 0; it < lsd_steps; it++) {
    // ---- ada = adaLN_all(silu(cproj + t_emb)); fx = input_proj(cur) ----
    LL_PREFETCH(P.ada_w, P.ada_n, C, nullptr, nullptr, 0);
    if (gw0 < P.ada_n || gw0 < C) {
      ll_stage_raw(sA, P.ll_cproj, C, t_cproj, pl);
      if (it == 0) { // the flow starts from the noise (an input of the launch)
        for (int i = threadIdx.x; i < ldim; i += FT)
          sA[C + i] = ldcg1(noise0 + i);
      } else {
        ll_stage_raw(sA + C, P.ll_cur, ldim, t_cur, pl);
      }
      for (int i0 = 2 * threadIdx.x; i0 < C; i0 += 2 * FT) { // own pairs of ll_stage_raw
#pragma unroll
        for (int e = 0; e < 2; e++) {
          const float y = sA[i0 + e] + __ldg(tvec + (long)it * C + i0 + e);
          sA[i0 + e] = y / (1.0f + expf(-y));
        }
      }
      __syncthreads();
      FSTAMP();
      gemv_cols(pf, pf2, P.ada_w, P.ada_n, C, sA, gw, nw, lane,
                [&](int n, float v) { ll_st(P.ll_ada, n, v + bias_at(P.ada_b, n), T); });
      gemv_now(P.ip_w, C, ldim, sA + C, gw, [&](int n, float v) { ll_st(P.ll_fx, n, v + bias_at(P.ip_b, n), T); });
    }
    const unsigned t_ada = T;
    unsigned t_fx = T;
    for (int b = 0; b < P.depth; b++) {
      const FusedRes &rb = P.rb[b];
      const uint2 *shift = P.ll_ada + (long)b * 3 * C, *scale = shift + C, *gate = shift + 2 * C;
      // ---- h2 = silu(mlp0(modulate(LN(fx)))) ----
      LL_PREFETCH(rb.mlp0w, P.hid, C, rb.lnw, rb.lnb, C);
      if (gw0 < P.hid) {
          ll_stage_ln(sA, red, P.ll_fx, C, t_fx, rb.lnw, rb.lnb, rb.eps, scale, shift, t_ada, pl);
        FSTAMP();
        gemv_cols(pf, pf2, rb.mlp0w, P.hid, C, sA, gw, nw, lane, [&](int n, float v) {
          ll_st(P.ll_h2, n, act_apply(v + bias_at(rb.mlp0b, n), ACT_SILU), T);
        });
      }
      const unsigned t_h2 = T;
      // ---- fx += gate * mlp2(h2) ----
      LL_PREFETCH(rb.mlp2w, C, P.hid, nullptr, nullptr, 0);
      if (gw0 < C) {
          ll_stage_plain(sA, P.ll_h2, P.hid, t_h2, pl);
        FSTAMP();
        float pre_g = 0.0f, pre_x = 0.0f;
        if (lane == 0 && gw < C) {
          pre_g = ll_get1(gate, gw, t_ada, pl);
          pre_x = ll_get1(P.ll_fx, gw, t_fx, pl);
        }
        gemv_cols(pf, pf2, rb.mlp2w, C, P.hid, sA, gw, nw, lane, [&](int n, float v) {
          const float g = n == gw ? pre_g : ll_get1(gate, n, t_ada, pl);
          const float xo = n == gw ? pre_x : ll_get1(P.ll_fx, n, t_fx, pl);
          ll_st(P.ll_fx, n, (v + bias_at(rb.mlp2b, n)) * g + xo, T);
        });
      }
      t_fx = T;
    }
    // ---- cur += final(modulate(LN(fx))) / steps ----
    const uint2 *fshift = P.ll_ada + (long)P.depth * 3 * C, *fscale = fshift + C;
    LL_PREFETCH(P.fin_w, ldim, C, nullptr, nullptr, 0);
    if (gw0 < ldim) {
      ll_stage_ln(sA, red, P.ll_fx, C, t_fx, nullptr, nullptr, 1e-6f, fscale, fshift, t_ada, pl);
      FSTAMP();
      const bool last_it = it == lsd_steps - 1;
      const bool first_it = it == 0;
      const unsigned t_prev = t_cur;
      gemv_cols(pf, pf2, P.fin_w, ldim, C, sA, gw, nw, lane, [&](int n, float v) {
        const float co = first_it ? ldcg1(noise0 + n) : ll_get1(P.ll_cur, n, t_prev, pl);
        const float c = (v + bias_at(P.fin_b, n)) * inv_steps + co;
        if (!last_it)
          ll_st(P.ll_cur, n, c, T);
        else {
          P.fh_cur[n] = c;
          P.hist_w[off_b[0] + n] = c; // the new latent frame of this sequence
        }
      });
    }
    t_cur = T;
  }
  FSTAMP();
  // Everybody that polls has read `epoch` long before CTA 0 gets here (CTA 0 consumed words
  // that depend on every producing CTA's first phase).
  if (blockIdx.x == 0 && threadIdx.x == 0)
    *P.epoch = T + 1;
#undef LL_PREFETCH
#undef L2_AHEAD
}

size_t step_fused_smem(const FusedParams &hp, int max_keys) {
  int n = hp.F;
  if (max_keys > n)
    n = max_keys;
  if (3 * hp.C > n) // row + adaLN scale + shift (flag-in-data variant)
    n = 3 * hp.C;
  if (2 * hp.C + hp.ldim > n) // activation staged twice for the row-pair GEMV + the latent
    n = 2 * hp.C + hp.ldim;
  return (size_t)((n + 15) / 16 * 16) * sizeof(float);
}

namespace {
inline size_t ev2(size_t n) { return (n + 1) & ~(size_t)1; } // 16-byte aligned sub-buffers
} // namespace
size_t step_fused_ll_words(const FusedParams &hp) {
  return ev2(hp.D) + ev2(3 * (size_t)hp.D) + ev2(hp.D) + ev2(hp.F) + ev2(hp.C) + ev2(hp.ada_n) + ev2(hp.C) +
         ev2(hp.hid) + ev2(hp.ldim);
}
void step_fused_ll_assign(FusedParams &hp, uint2 *p) {
  hp.ll_x = p;      p += ev2(hp.D);
  hp.ll_qkv = p;    p += ev2(3 * (size_t)hp.D);
  hp.ll_att = p;    p += ev2(hp.D);
  hp.ll_ff = p;     p += ev2(hp.F);
  hp.ll_cproj = p;  p += ev2(hp.C);
  hp.ll_ada = p;    p += ev2(hp.ada_n);
  hp.ll_fx = p;     p += ev2(hp.C);
  hp.ll_h2 = p;     p += ev2(hp.hid);
  hp.ll_cur = p;
}

int step_fused_grid(size_t smem, bool ll) {
  int dev = 0, sms = 0, occ = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const void *fn = ll ? (const void *)step_fused_ll_kernel : (const void *)step_fused_kernel;
  if (cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
    return 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fn, FT, smem) != cudaSuccess || occ < 1)
    return 0;
  if (occ > 2)
    occ = 2;
  return occ * sms;
}

int step_fused_launch(const FusedParams *dp, int grid, size_t smem, const int32_t *row_slot, const int32_t *row_pos,
                      const int64_t *off_a, const int64_t *off_b, const int64_t *off_c, int gather_frames,
                      int gather_noise, const float *tvec, int lsd_steps, unsigned *barrier, bool ll, cudaStream_t st) {
  if (!ll && cudaMemsetAsync(barrier, 0, STEP_FUSED_BARRIER_BYTES, st) != cudaSuccess)
    return -1;
  void *args[] = {(void *)&dp,       (void *)&row_slot,      (void *)&row_pos,      (void *)&off_a, (void *)&off_b,
                  (void *)&off_c,    (void *)&gather_frames, (void *)&gather_noise, (void *)&tvec,  (void *)&lsd_steps};
  // cooperative in both variants: co-residency of all CTAs is what makes the spinning legal
  void *fn = ll ? (void *)step_fused_ll_kernel : (void *)step_fused_kernel;
  if (cudaLaunchCooperativeKernel(fn, dim3(grid), dim3(FT), args, smem, st) != cudaSuccess)
    return -1;
  return ll ? 1 : 2; // (memset +) kernel
}

} // namespace ptts
