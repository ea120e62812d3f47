// attention.cu — RoPE + paged KV append and the two masked-attention kernels.
//
// Reference semantics (internal/runtime/ops/attention.go:307-484): per (row,
// head) scores s_k = (q.k_k)/sqrt(Dh) over the allowed keys only (masked keys
// are SKIPPED, never multiplied: unused slots of upstream voice files may hold
// NaN, scripts/export_onnx.py:114), max, exp, sum, out = sum_k (e_k/sum) v_k;
// a row with no allowed key yields zeros.  The reference evaluates exp and the
// sum in fp64; here they are fp32 (expf, <= 8192 positive terms), which is far
// inside the 2e-4 attention tolerance of ops/tolerance.go:13-24.
//
// One CTA of 128 threads per (row, head), head_dim == 64:
//   phase 1: 16 keys per pass (8 lanes x 2 float4 per key, 3 shuffles) -> smem
//   phase 2: block max / exp / sum
//   phase 3: P.V with 16 float4 columns x 8 key groups, smem cross-group reduce
// A KV page (16 tokens x 64 floats for one head) is one contiguous 4 KB run, so
// phase 1 and 3 read whole 128-byte lines.
#include "kernels.cuh"
#include "split.cuh"
#include "launch.cuh"
#include "attend.cuh"
#include <math.h>
#include <stdlib.h>

namespace ptts {

template <typename KT, bool POW2>
__global__ void __launch_bounds__(ATT_THREADS)
    attention_paged_kernel(const float *qkv, long ldq, float *out, long ldo,
                           const int32_t *row_slot, const int32_t *row_pos,
                           KvGeom kv, int layer, int pt_shift, SplitOut sp) {
  PDL_TRIGGER();
  PDL_WAIT();
  extern __shared__ float smem[];
  __shared__ float red[ATT_THREADS / 32];
  __shared__ float4 pv[8 * 16];
  const int r = blockIdx.x, h = blockIdx.y;
  const int slot = row_slot[r], pos = row_pos[r];
  float *o = out ? out + (long)r * ldo + h * HD : nullptr;
  if (pos < 0) {
    if (threadIdx.x < 16) {
      const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
      if (o)
        *reinterpret_cast<float4 *>(o + threadIdx.x * 4) = z;
      if (sp.hi)
        split_store4(sp, r, h * HD + threadIdx.x * 4, z);
    }
    return;
  }
  const int nkeys = pos + 1;
  const int32_t *pt = kv.pagetab + (long)slot * kv.max_pages;
  const int PT = kv.page_tokens;
  const long page_stride = (long)kv.n_layers * 2 * kv.n_heads * PT * HD;
  const long koff = ((long)(layer * 2 + 0) * kv.n_heads + h) * PT * HD;
  const long voff = ((long)(layer * 2 + 1) * kv.n_heads + h) * PT * HD;
  const KT *pool = static_cast<const KT *>(kv.pool);
  const int pmask = PT - 1; // POW2: PT == 1 << pt_shift (see attention_paged_decode_kernel)
  auto keyp = [&](int j) {
    if (POW2)
      return pool + (long)pt[j >> pt_shift] * page_stride + koff + ((j & pmask) << 6);
    return pool + (long)pt[j / PT] * page_stride + koff + (long)(j % PT) * HD;
  };
  auto valp = [&](int j) {
    if (POW2)
      return pool + (long)pt[j >> pt_shift] * page_stride + voff + ((j & pmask) << 6);
    return pool + (long)pt[j / PT] * page_stride + voff + (long)(j % PT) * HD;
  };
  attend_row(qkv + (long)r * ldq + h * HD, nkeys, keyp, valp, o, smem, red, pv, sp, r, h * HD);
}

// Decode step (one new row per sequence): RoPE of q and k (rope.go:81-105), the K/V append of
// the row's own token (flow_transformer.go:32-67) and the attention over positions 0..pos in ONE
// kernel.  The CTA of (row, head) rotates its q into shared memory, writes the rotated k and the v
// of its head into the sequence's page, and then attends over the cache, its own token included.
// Only valid when no other row of the launch belongs to the same sequence (the prefill keeps the
// separate append kernel: there a row also reads keys written by other CTAs).
// POW2: the page size is a power of two, 1 << pt_shift (key -> page / row by shift and mask; the integer
// divisions by a run-time page size were a third of the instructions this issue-bound kernel executes).
// Pages of exactly 16 tokens (the default): one (row, head) of decode attention with the loops written per page —
// a score pass is one page (one page-table lookup, no per-key guard), a value step two pages (four rows in
// flight); the partial last page is the only guarded part.  Same thread-to-element assignment, summation order
// and softmax as attend_row (attend.cuh): bit-identical output.  The kernel is issue-bound, so what this removes
// (address updates, guards, page lookups per key) is time.
template <typename KT>
__device__ __forceinline__ void attend_row_pages16(const float *q, int nkeys, const KT *kbase, const KT *vbase,
                                                   const int32_t *s_pt, long page_stride, float *out, float *scores,
                                                   float *red, float4 *pv, const SplitOut &sp, long sp_row, int sp_col) {
  const int tid = threadIdx.x;
  const float scale = 0.125f; // 1/sqrt(64)
  const int npg = nkeys >> 4, rem = nkeys & 15;
  // ---- scores ----
  const int part = tid & 7, kslot = tid >> 3;
  const float4 q0 = *reinterpret_cast<const float4 *>(q + part * 8);
  const float4 q1 = *reinterpret_cast<const float4 *>(q + part * 8 + 4);
  const KT *krow = kbase + kslot * HD + part * 8;
  auto dot8 = [&](const KT *p) {
    float k[8];
    kv_load8(p, k);
    float s = q0.x * k[0];
    s = fmaf(q0.y, k[1], s);
    s = fmaf(q0.z, k[2], s);
    s = fmaf(q0.w, k[3], s);
    s = fmaf(q1.x, k[4], s);
    s = fmaf(q1.y, k[5], s);
    s = fmaf(q1.z, k[6], s);
    s = fmaf(q1.w, k[7], s);
    return s;
  };
  float lmax = -INFINITY;
#pragma unroll 2
  for (int p = 0; p < npg; p++) {
    float s = dot8(krow + (long)s_pt[p] * page_stride);
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    s += __shfl_xor_sync(0xffffffffu, s, 4);
    s *= scale;
    if (part == 0)
      scores[p * 16 + kslot] = s;
    lmax = fmaxf(lmax, s);
  }
  if (rem) {
    const bool ok = kslot < rem;
    float s = ok ? dot8(krow + (long)s_pt[npg] * page_stride) : 0.0f;
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    s += __shfl_xor_sync(0xffffffffu, s, 4);
    if (ok) {
      s *= scale;
      if (part == 0)
        scores[npg * 16 + kslot] = s;
      lmax = fmaxf(lmax, s);
    }
  }
  const float gmax = block_reduce_max(lmax, red); // includes __syncthreads
  // ---- exp / sum ----
  float lsum = 0.0f;
  for (int j = tid; j < nkeys; j += ATT_THREADS) {
    float e = expf(scores[j] - gmax);
    scores[j] = e;
    lsum += e;
  }
  const float gsum = block_reduce_sum(lsum, red);
  const float inv = 1.0f / gsum;
  // ---- P.V: thread (grp, c4) owns rows grp and grp + 8 of every page, ascending ----
  const int c4 = tid & 15, grp = tid >> 4;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  const KT *vrow = vbase + grp * HD + c4 * 4;
  auto fma4 = [&](float w, const float4 &v) {
    acc.x = fmaf(w, v.x, acc.x);
    acc.y = fmaf(w, v.y, acc.y);
    acc.z = fmaf(w, v.z, acc.z);
    acc.w = fmaf(w, v.w, acc.w);
  };
  int p = 0;
  for (; p + 1 < npg; p += 2) {
    const KT *a = vrow + (long)s_pt[p] * page_stride, *b = vrow + (long)s_pt[p + 1] * page_stride;
    const float4 v0 = kv_load4(a), v1 = kv_load4(a + 8 * HD), v2 = kv_load4(b), v3 = kv_load4(b + 8 * HD);
    const float *sc = scores + p * 16 + grp;
    const float w0 = sc[0] * inv, w1 = sc[8] * inv, w2 = sc[16] * inv, w3 = sc[24] * inv;
    fma4(w0, v0);
    fma4(w1, v1);
    fma4(w2, v2);
    fma4(w3, v3);
  }
  if (p < npg) {
    const KT *a = vrow + (long)s_pt[p] * page_stride;
    const float4 v0 = kv_load4(a), v1 = kv_load4(a + 8 * HD);
    const float w0 = scores[p * 16 + grp] * inv, w1 = scores[p * 16 + grp + 8] * inv;
    fma4(w0, v0);
    fma4(w1, v1);
  }
  if (rem) {
    const KT *a = vrow + (long)s_pt[npg] * page_stride;
    if (grp < rem)
      fma4(scores[npg * 16 + grp] * inv, kv_load4(a));
    if (grp + 8 < rem)
      fma4(scores[npg * 16 + grp + 8] * inv, kv_load4(a + 8 * HD));
  }
  pv[grp * 16 + c4] = acc;
  __syncthreads();
  if (tid < 16) {
    float4 r = pv[tid];
#pragma unroll
    for (int g2 = 1; g2 < 8; g2++) {
      float4 t = pv[g2 * 16 + tid];
      r.x += t.x; r.y += t.y; r.z += t.z; r.w += t.w;
    }
    if (out)
      *reinterpret_cast<float4 *>(out + tid * 4) = r;
    if (sp.hi)
      split_store4(sp, sp_row, sp_col + tid * 4, r);
  }
}

// MODE 0: any page size (integer division); 1: power-of-two pages (shift / mask); 2: 16-token pages, page-wise loops;
// 3: 16-token pages, one-pass flash decoding (attend_row_flash16).
template <typename KT, int MODE>
__global__ void __launch_bounds__(ATT_THREADS)
    attention_paged_decode_kernel(const float *qkv, long ldq, float *out, long ldo,
                                  const int32_t *row_slot, const int32_t *row_pos, const float *cosT,
                                  const float *sinT, KvGeom kv, int layer, int pt_shift, SplitOut sp) {
  // Everything up to PDL_WAIT reads only what no kernel of the step writes (row tables, the page table, the RoPE
  // tables) and writes nothing global: it overlaps the tail of the in_proj GEMM this launch depends on.
  PDL_TRIGGER();
  extern __shared__ float smem[];
  __shared__ float red[ATT_THREADS / 32];
  __shared__ float4 pv[8 * 16];
  __shared__ __align__(16) float qs[HD];
  const int r = blockIdx.x, h = blockIdx.y, tid = threadIdx.x;
  const int slot = row_slot[r], pos = row_pos[r];
  float *o = out ? out + (long)r * ldo + h * HD : nullptr;
  if (pos < 0) {
    PDL_WAIT();
    if (tid < 16) {
      const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
      if (o)
        *reinterpret_cast<float4 *>(o + tid * 4) = z;
      if (sp.hi)
        split_store4(sp, r, h * HD + tid * 4, z);
    }
    return;
  }
  const int nkeys = pos + 1;
  const int32_t *pt = kv.pagetab + (long)slot * kv.max_pages;
  const int PT = kv.page_tokens, D = kv.n_heads * HD;
  const long page_stride = (long)kv.n_layers * 2 * kv.n_heads * PT * HD;
  const long koff = ((long)(layer * 2 + 0) * kv.n_heads + h) * PT * HD;
  const long voff = ((long)(layer * 2 + 1) * kv.n_heads + h) * PT * HD;
  KT *pool = static_cast<KT *>(kv.pool);
  // the sequence's page ids go to shared memory once: every key / value access below would otherwise be two
  // dependent global loads (page id, then the row)
  __shared__ int32_t s_pt[KV_MAX_SEQ_PAGES];
  for (int i = tid; i * PT < nkeys; i += ATT_THREADS)
    s_pt[i] = pt[i];
  const int pmask = PT - 1;
  float cj = 0.0f, sj = 0.0f;
  if (tid < HD) {
    cj = cosT[(long)pos * (HD / 2) + (tid & (HD / 2 - 1))];
    sj = sinT[(long)pos * (HD / 2) + (tid & (HD / 2 - 1))];
  }
  PDL_WAIT();
  {
    const float *row = qkv + (long)r * ldq;
    KT *page = pool + (long)pt[MODE ? pos >> pt_shift : pos / PT] * page_stride +
               (long)(MODE ? pos & pmask : pos % PT) * HD;
    if (tid < HD) { // pairs (2j, 2j+1): threads 0..31 rotate q, 32..63 rotate k
      const int isk = tid >= HD / 2, j = tid & (HD / 2 - 1);
      const float *src = row + (isk ? D : 0) + h * HD + 2 * j;
      const float a = src[0], b = src[1];
      const float o0 = rope_even(a, b, cj, sj);
      const float o1 = rope_odd(a, b, cj, sj);
      if (isk) {
        kv_store(page + koff + 2 * j, o0);
        kv_store(page + koff + 2 * j + 1, o1);
      } else {
        qs[2 * j] = o0;
        qs[2 * j + 1] = o1;
      }
    } else { // threads 64..127: the head's v
      const int d = tid - HD;
      kv_store(page + voff + d, row[2 * D + h * HD + d]);
    }
  }
  __syncthreads(); // q and the page ids in shared memory; this CTA's own k/v stores are visible to its loads below
  const KT *kbase = pool + koff, *vbase = pool + voff;
  if (MODE == 2) {
    attend_row_pages16(qs, nkeys, kbase, vbase, s_pt, page_stride, o, smem, red, pv, sp, r, h * HD);
    return;
  }
  if (MODE == 3) {
    attend_row_flash16(qs, nkeys, kbase, vbase, s_pt, page_stride, o, smem, sp, r, h * HD);
    return;
  }
  auto keyp = [&](int j) {
    if (MODE)
      return kbase + (long)s_pt[j >> pt_shift] * page_stride + ((j & pmask) << 6);
    return kbase + (long)s_pt[j / PT] * page_stride + (long)(j % PT) * HD;
  };
  auto valp = [&](int j) {
    if (MODE)
      return vbase + (long)s_pt[j >> pt_shift] * page_stride + ((j & pmask) << 6);
    return vbase + (long)s_pt[j / PT] * page_stride + (long)(j % PT) * HD;
  };
  attend_row(qs, nkeys, keyp, valp, o, smem, red, pv, sp, r, h * HD);
}

// Two heads of one row per CTA, 64 threads per head (16-token pages; opt-in, see attention_paged_decode).  The
// kernel above is issue-bound and more than half of what it issues at kv ~ 100 is per-CTA overhead that every
// warp repeats (row / page-table setup, block reductions, barriers): here four warps carry two heads, so that
// overhead is paid once per two heads.  Measured: no faster.
// Each thread plays two threads of the 128-thread layout — keys kslot and kslot + 8 of a page in the score pass,
// the partial sums of threads t and t + 64 in the softmax denominator, value groups grp and grp + 4 — and every
// sum is formed in the same order, so the output is bit-identical to attend_row / attend_row_pages16.
template <typename KT>
__global__ void __launch_bounds__(ATT_THREADS)
    attention_paged_decode2_kernel(const float *qkv, long ldq, float *out, long ldo, const int32_t *row_slot,
                                   const int32_t *row_pos, const float *cosT, const float *sinT, KvGeom kv, int layer,
                                   int score_floats, SplitOut sp) {
  PDL_TRIGGER();
  extern __shared__ float smem[]; // scores of the two heads
  __shared__ float red[2][4];
  __shared__ float4 pv[2][8 * 16];
  __shared__ __align__(16) float qs2[2][HD];
  __shared__ int32_t s_pt[KV_MAX_SEQ_PAGES];
  const int r = blockIdx.x, tid = threadIdx.x, sub = tid >> 6, t = tid & 63, lane = tid & 31, w = t >> 5;
  const int h = blockIdx.y * 2 + sub;
  const int slot = row_slot[r], pos = row_pos[r];
  float *o = out ? out + (long)r * ldo + h * HD : nullptr;
  if (pos < 0) {
    PDL_WAIT();
    if (t < 16) {
      const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
      if (o)
        *reinterpret_cast<float4 *>(o + t * 4) = z;
      if (sp.hi)
        split_store4(sp, r, h * HD + t * 4, z);
    }
    return;
  }
  const int nkeys = pos + 1, npg = nkeys >> 4, rem = nkeys & 15;
  const int32_t *pt = kv.pagetab + (long)slot * kv.max_pages;
  const int D = kv.n_heads * HD;
  const long page_stride = (long)kv.n_layers * 2 * kv.n_heads * 16 * HD;
  const long koff = ((long)(layer * 2 + 0) * kv.n_heads + h) * 16 * HD;
  const long voff = ((long)(layer * 2 + 1) * kv.n_heads + h) * 16 * HD;
  KT *pool = static_cast<KT *>(kv.pool);
  for (int i = tid; (i << 4) < nkeys; i += ATT_THREADS)
    s_pt[i] = pt[i];
  float cj = 0.0f, sj = 0.0f;
  {
    const int j = t & (HD / 2 - 1);
    cj = cosT[(long)pos * (HD / 2) + j];
    sj = sinT[(long)pos * (HD / 2) + j];
  }
  float *qs = qs2[sub];
  float *scores = smem + sub * score_floats;
  PDL_WAIT();
  {
    const float *row = qkv + (long)r * ldq;
    KT *page = pool + (long)pt[pos >> 4] * page_stride + (long)(pos & 15) * HD;
    const int isk = t >= HD / 2, j = t & (HD / 2 - 1); // threads 0..31 of the head rotate q, 32..63 rotate k
    const float *src = row + (isk ? D : 0) + h * HD + 2 * j;
    const float a = src[0], b = src[1];
    const float o0 = rope_even(a, b, cj, sj);
    const float o1 = rope_odd(a, b, cj, sj);
    if (isk) {
      kv_store(page + koff + 2 * j, o0);
      kv_store(page + koff + 2 * j + 1, o1);
    } else {
      qs[2 * j] = o0;
      qs[2 * j + 1] = o1;
    }
    kv_store(page + voff + t, row[2 * D + h * HD + t]); // and every thread one element of the head's v
  }
  __syncthreads(); // q and the page ids in shared memory; this CTA's own k/v stores are visible to its loads below
  const KT *kbase = pool + koff, *vbase = pool + voff;
  const float scale = 0.125f; // 1/sqrt(64)
  // ---- scores: rows kslot and kslot + 8 of every page ----
  const int part = t & 7, kslot = t >> 3; // 0..7
  const float4 q0 = *reinterpret_cast<const float4 *>(qs + part * 8);
  const float4 q1 = *reinterpret_cast<const float4 *>(qs + part * 8 + 4);
  const KT *krow = kbase + kslot * HD + part * 8;
  auto dot8 = [&](const KT *p) {
    float k[8];
    kv_load8(p, k);
    float s = q0.x * k[0];
    s = fmaf(q0.y, k[1], s);
    s = fmaf(q0.z, k[2], s);
    s = fmaf(q0.w, k[3], s);
    s = fmaf(q1.x, k[4], s);
    s = fmaf(q1.y, k[5], s);
    s = fmaf(q1.z, k[6], s);
    s = fmaf(q1.w, k[7], s);
    return s;
  };
  auto red8 = [&](float s) {
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    s += __shfl_xor_sync(0xffffffffu, s, 4);
    return s;
  };
  float lmax = -INFINITY;
  for (int p = 0; p < npg; p++) {
    const KT *kp = krow + (long)s_pt[p] * page_stride;
    float s0 = dot8(kp), s1 = dot8(kp + 8 * HD);
    s0 = red8(s0) * scale;
    s1 = red8(s1) * scale;
    if (part == 0) {
      scores[p * 16 + kslot] = s0;
      scores[p * 16 + kslot + 8] = s1;
    }
    lmax = fmaxf(lmax, fmaxf(s0, s1));
  }
  if (rem) {
    const KT *kp = krow + (long)s_pt[npg] * page_stride;
    const bool ok0 = kslot < rem, ok1 = kslot + 8 < rem;
    float s0 = ok0 ? dot8(kp) : 0.0f, s1 = ok1 ? dot8(kp + 8 * HD) : 0.0f;
    s0 = red8(s0);
    s1 = red8(s1);
    if (ok0) {
      s0 *= scale;
      if (part == 0)
        scores[npg * 16 + kslot] = s0;
      lmax = fmaxf(lmax, s0);
    }
    if (ok1) {
      s1 *= scale;
      if (part == 0)
        scores[npg * 16 + kslot + 8] = s1;
      lmax = fmaxf(lmax, s1);
    }
  }
  // max over the head's 64 threads
#pragma unroll
  for (int of = 16; of > 0; of >>= 1)
    lmax = fmaxf(lmax, __shfl_xor_sync(0xffffffffu, lmax, of));
  if (lane == 0)
    red[sub][w] = lmax;
  __syncthreads();
  const float gmax = fmaxf(red[sub][0], red[sub][1]);
  __syncthreads();
  // ---- exp / sum: the partial sums of threads t and t + 64 of the 128-thread layout, reduced per (virtual) warp ----
  float la = 0.0f, lb = 0.0f;
  for (int j = t; j < nkeys; j += ATT_THREADS) {
    const float e = expf(scores[j] - gmax);
    scores[j] = e;
    la += e;
    if (j + 64 < nkeys) {
      const float e2 = expf(scores[j + 64] - gmax);
      scores[j + 64] = e2;
      lb += e2;
    }
  }
#pragma unroll
  for (int of = 16; of > 0; of >>= 1) {
    la += __shfl_xor_sync(0xffffffffu, la, of);
    lb += __shfl_xor_sync(0xffffffffu, lb, of);
  }
  if (lane == 0) { // virtual warps w (threads t) and w + 2 (threads t + 64)
    red[sub][w] = la;
    red[sub][w + 2] = lb;
  }
  __syncthreads();
  float gsum = 0.0f;
#pragma unroll
  for (int i = 0; i < 4; i++)
    gsum += red[sub][i];
  const float inv = 1.0f / gsum;
  // ---- P.V: value groups grp and grp + 4 (rows grp, grp + 8 and grp + 4, grp + 12 of every page), ascending ----
  const int c4 = t & 15, grp = t >> 4; // 0..3
  float4 acA = make_float4(0.f, 0.f, 0.f, 0.f), acB = acA;
  const KT *vrow = vbase + grp * HD + c4 * 4;
  auto fma4 = [](float4 &acc, float wv, const float4 &v) {
    acc.x = fmaf(wv, v.x, acc.x);
    acc.y = fmaf(wv, v.y, acc.y);
    acc.z = fmaf(wv, v.z, acc.z);
    acc.w = fmaf(wv, v.w, acc.w);
  };
  for (int p = 0; p < npg; p++) {
    const KT *a = vrow + (long)s_pt[p] * page_stride;
    const float4 v0 = kv_load4(a), v1 = kv_load4(a + 4 * HD), v2 = kv_load4(a + 8 * HD), v3 = kv_load4(a + 12 * HD);
    const float *sc = scores + p * 16 + grp;
    const float w0 = sc[0] * inv, w1 = sc[4] * inv, w2 = sc[8] * inv, w3 = sc[12] * inv;
    fma4(acA, w0, v0);
    fma4(acB, w1, v1);
    fma4(acA, w2, v2);
    fma4(acB, w3, v3);
  }
  if (rem) {
    const KT *a = vrow + (long)s_pt[npg] * page_stride;
    const float *sc = scores + npg * 16 + grp;
    if (grp < rem)
      fma4(acA, sc[0] * inv, kv_load4(a));
    if (grp + 4 < rem)
      fma4(acB, sc[4] * inv, kv_load4(a + 4 * HD));
    if (grp + 8 < rem)
      fma4(acA, sc[8] * inv, kv_load4(a + 8 * HD));
    if (grp + 12 < rem)
      fma4(acB, sc[12] * inv, kv_load4(a + 12 * HD));
  }
  pv[sub][grp * 16 + c4] = acA;
  pv[sub][(grp + 4) * 16 + c4] = acB;
  __syncthreads();
  if (t < 16) {
    float4 rr = pv[sub][t];
#pragma unroll
    for (int g2 = 1; g2 < 8; g2++) {
      const float4 t4 = pv[sub][g2 * 16 + t];
      rr.x += t4.x; rr.y += t4.y; rr.z += t4.z; rr.w += t4.w;
    }
    if (o)
      *reinterpret_cast<float4 *>(o + t * 4) = rr;
    if (sp.hi)
      split_store4(sp, r, h * HD + t * 4, rr);
  }
}

__global__ void __launch_bounds__(ATT_THREADS)
    attention_window_kernel(const float *qkv, long ldq, float *out, long ldo,
                            int D, const int32_t *row_pos, int ctx, SplitOut sp) {
  extern __shared__ float smem[];
  __shared__ float red[ATT_THREADS / 32];
  __shared__ float4 pv[8 * 16];
  const int r = blockIdx.x, h = blockIdx.y;
  const int pos = row_pos[r];
  float *o = out ? out + (long)r * ldo + h * HD : nullptr;
  if (pos < 0) {
    if (threadIdx.x < 16) {
      const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
      if (o)
        *reinterpret_cast<float4 *>(o + threadIdx.x * 4) = z;
      if (sp.hi)
        split_store4(sp, r, h * HD + threadIdx.x * 4, z);
    }
    return;
  }
  // allowed keys: 0 <= pos - pk < ctx  ->  the last min(ctx, pos+1) rows
  const int nkeys = (ctx > 0 && pos + 1 > ctx) ? ctx : pos + 1;
  const int first = r - (nkeys - 1);
  const float *kbase = qkv + D + h * HD;
  const float *vbase = qkv + 2 * D + h * HD;
  auto keyp = [&](int j) { return kbase + (long)(first + j) * ldq; };
  auto valp = [&](int j) { return vbase + (long)(first + j) * ldq; };
  attend_row(qkv + (long)r * ldq + h * HD, nkeys, keyp, valp, o, smem, red, pv, sp, r, h * HD);
}

// Multi-query windowed attention: MQ consecutive rows per CTA share every K/V
// load (their windows overlap by ctx-1 rows), cutting the L2 traffic of the
// one-row-per-CTA kernel ~MQ-fold.  Same arithmetic per (query, key) pair.
constexpr int MQ = 8;
__global__ void __launch_bounds__(ATT_THREADS)
    attention_window_mq_kernel(const float *qkv, long ldq, float *out, long ldo, int rows, int D,
                               const int32_t *row_pos, int ctx, SplitOut sp) {
  extern __shared__ float smem[];          // scores [MQ][span_cap]
  __shared__ __align__(16) float qs[MQ][HD];
  __shared__ float4 pv[8][MQ][16];
  __shared__ int s_lo[MQ], s_n[MQ];
  __shared__ float s_inv[MQ];
  const int tid = threadIdx.x, h = blockIdx.y;
  const int r0 = blockIdx.x * MQ;
  const int span_cap = ctx + MQ;
  if (tid < MQ) {
    const int r = r0 + tid;
    int n = 0, lo = r;
    if (r < rows) {
      const int pos = row_pos[r];
      if (pos >= 0) {
        n = (ctx > 0 && pos + 1 > ctx) ? ctx : pos + 1;
        lo = r - (n - 1);
      }
    }
    s_lo[tid] = lo;
    s_n[tid] = n;
  }
  for (int i = tid; i < MQ * HD; i += ATT_THREADS) {
    const int qi = i / HD, d = i % HD;
    const int r = r0 + qi;
    qs[qi][d] = r < rows ? qkv[(long)r * ldq + h * HD + d] : 0.0f;
  }
  __syncthreads();
  int lo_min = r0 + MQ, hi = r0 - 1;
#pragma unroll
  for (int i = 0; i < MQ; i++)
    if (s_n[i] > 0) {
      lo_min = min(lo_min, s_lo[i]);
      hi = max(hi, r0 + i);
    }
  const int span = hi - lo_min + 1; // <= ctx + MQ - 1
  const float *kbase = qkv + D + h * HD, *vbase = qkv + 2 * D + h * HD;
  // phase 1: scores[q][jj] for every key row of the tile's span
  {
    const int part = tid & 7, kslot = tid >> 3;
    const int my_lo = s_lo[part], my_hi = r0 + part, my_n = s_n[part];
    for (int j0 = 0; j0 < span; j0 += 16) {
      const int jj = j0 + kslot;
      float sc[MQ];
#pragma unroll
      for (int i = 0; i < MQ; i++)
        sc[i] = 0.0f;
      if (jj < span) {
        const float *kp = kbase + (long)(lo_min + jj) * ldq;
        const float4 a = *reinterpret_cast<const float4 *>(kp + part * 4);
        const float4 b = *reinterpret_cast<const float4 *>(kp + 32 + part * 4);
#pragma unroll
        for (int i = 0; i < MQ; i++) {
          const float4 q0 = *reinterpret_cast<const float4 *>(&qs[i][part * 4]);
          const float4 q1 = *reinterpret_cast<const float4 *>(&qs[i][32 + part * 4]);
          float t = q0.x * a.x;
          t = fmaf(q0.y, a.y, t); t = fmaf(q0.z, a.z, t); t = fmaf(q0.w, a.w, t);
          t = fmaf(q1.x, b.x, t); t = fmaf(q1.y, b.y, t); t = fmaf(q1.z, b.z, t); t = fmaf(q1.w, b.w, t);
          sc[i] = t;
        }
      }
#pragma unroll
      for (int i = 0; i < MQ; i++) {
        float t = sc[i];
        t += __shfl_xor_sync(0xffffffffu, t, 1);
        t += __shfl_xor_sync(0xffffffffu, t, 2);
        t += __shfl_xor_sync(0xffffffffu, t, 4);
        sc[i] = t;
      }
      if (jj < span) {
        // lane `part` publishes the score of query `part`
        float mine = 0.0f;
#pragma unroll
        for (int i = 0; i < MQ; i++)
          if (i == part)
            mine = sc[i];
        const int j = lo_min + jj;
        const bool ok = my_n > 0 && j >= my_lo && j <= my_hi;
        smem[part * span_cap + jj] = ok ? mine * 0.125f : -INFINITY;
      }
    }
  }
  __syncthreads();
  // phase 2: per-query max / exp / sum (one warp handles MQ/4 queries)
  {
    const int warp = tid >> 5, lane = tid & 31;
    for (int qi = warp; qi < MQ; qi += ATT_THREADS / 32) {
      float *sq = smem + qi * span_cap;
      float m = -INFINITY;
      for (int jj = lane; jj < span; jj += 32)
        m = fmaxf(m, sq[jj]);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1)
        m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
      float sum = 0.0f;
      for (int jj = lane; jj < span; jj += 32) {
        const float v = sq[jj];
        const float e = v == -INFINITY ? 0.0f : expf(v - m);
        sq[jj] = e;
        sum += e;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1)
        sum += __shfl_xor_sync(0xffffffffu, sum, o);
      if (lane == 0)
        s_inv[qi] = sum > 0.0f ? 1.0f / sum : 0.0f;
    }
  }
  __syncthreads();
  // phase 3: out[q] = sum_j p[q][j] v_j ; V row loaded once for all MQ queries
  {
    const int c4 = tid & 15, grp = tid >> 4;
    float4 acc[MQ];
#pragma unroll
    for (int i = 0; i < MQ; i++)
      acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int jj = grp; jj < span; jj += 8) {
      const float4 v = *reinterpret_cast<const float4 *>(vbase + (long)(lo_min + jj) * ldq + c4 * 4);
#pragma unroll
      for (int i = 0; i < MQ; i++) {
        const float w = smem[i * span_cap + jj] * s_inv[i];
        if (w != 0.0f) { // masked keys are skipped, never multiplied (attention.go:403-406)
          acc[i].x = fmaf(w, v.x, acc[i].x);
          acc[i].y = fmaf(w, v.y, acc[i].y);
          acc[i].z = fmaf(w, v.z, acc[i].z);
          acc[i].w = fmaf(w, v.w, acc[i].w);
        }
      }
    }
#pragma unroll
    for (int i = 0; i < MQ; i++)
      pv[grp][i][c4] = acc[i];
  }
  __syncthreads();
  {
    const int qi = tid >> 4, c4 = tid & 15; // 8 queries x 16 float4 columns = 128 threads
    const int r = r0 + qi;
    if (r < rows) {
      float4 t = pv[0][qi][c4];
#pragma unroll
      for (int g2 = 1; g2 < 8; g2++) {
        const float4 u = pv[g2][qi][c4];
        t.x += u.x; t.y += u.y; t.z += u.z; t.w += u.w;
      }
      if (out)
        *reinterpret_cast<float4 *>(out + (long)r * ldo + h * HD + c4 * 4) = t;
      if (sp.hi)
        split_store4(sp, r, h * HD + c4 * 4, t);
    }
  }
}

// log2 of a power-of-two page size, -1 otherwise
static int page_shift(int PT) {
  if (PT <= 0 || (PT & (PT - 1)) != 0)
    return -1;
  int s = 0;
  while ((1 << s) < PT)
    s++;
  return s;
}

int attention_paged(const float *qkv, long ldq, float *out, long ldo, int rows,
                    const int32_t *row_slot, const int32_t *row_pos,
                    const KvGeom &kv, int layer, int max_keys, cudaStream_t st, SplitOut sp) {
  if (rows <= 0)
    return 0;
  dim3 grid(rows, kv.n_heads);
  size_t sm = sizeof(float) * (size_t)((max_keys + 15) / 16 * 16);
  const int pt_shift = page_shift(kv.page_tokens);
  if (kv.half && pt_shift >= 0)
    launch_k(attention_paged_kernel<__half, true>, grid, dim3(ATT_THREADS), sm, st, qkv, ldq, out, ldo, row_slot, row_pos,
             kv, layer, pt_shift, sp);
  else if (kv.half)
    launch_k(attention_paged_kernel<__half, false>, grid, dim3(ATT_THREADS), sm, st, qkv, ldq, out, ldo, row_slot,
             row_pos, kv, layer, 0, sp);
  else if (pt_shift >= 0)
    launch_k(attention_paged_kernel<float, true>, grid, dim3(ATT_THREADS), sm, st, qkv, ldq, out, ldo, row_slot, row_pos,
             kv, layer, pt_shift, sp);
  else
    launch_k(attention_paged_kernel<float, false>, grid, dim3(ATT_THREADS), sm, st, qkv, ldq, out, ldo, row_slot, row_pos,
             kv, layer, 0, sp);
  return 1;
}

int attention_paged_decode(const float *qkv, long ldq, float *out, long ldo, int rows, const int32_t *row_slot,
                           const int32_t *row_pos, const float *cosT, const float *sinT, const KvGeom &kv, int layer,
                           int max_keys, cudaStream_t st, SplitOut sp) {
  if (rows <= 0)
    return 0;
  // opt-in: K / V staged through shared memory by cp.async (attention_decode.cu).  Bit-identical; per launch it
  // sits between the old and the new register kernel (64 vs 69 / 57 us at B = 512) and its 34 KB of shared
  // memory per CTA keep the next GEMM's CTAs off the SM until the last attention CTA has left.
  if (kv.half && getenv("PTTS_ATTN_STAGED") != nullptr) {
    const int n = attention_paged_decode_staged(qkv, ldq, out, ldo, rows, row_slot, row_pos, cosT, sinT, kv, layer,
                                                max_keys, st, sp);
    if (n > 0)
      return n;
  }
  dim3 grid(rows, kv.n_heads);
  size_t sm = sizeof(float) * (size_t)((max_keys + 15) / 16 * 16);
  const int pt_shift = page_shift(kv.page_tokens);
#define PTTS_DECODE_LAUNCH(KT, MODE)                                                                              \
  launch_k(attention_paged_decode_kernel<KT, MODE>, grid, dim3(ATT_THREADS), sm, st, qkv, ldq, out, ldo, row_slot,   \
           row_pos, cosT, sinT, kv, layer, pt_shift < 0 ? 0 : pt_shift, sp)
  int mode = pt_shift == 4 && getenv("PTTS_ATTN_NO_PAGEWISE") == nullptr ? 2 : (pt_shift >= 0 ? 1 : 0);
  {
    // one-pass flash decoding for batches too small to hide a CTA's latency chain (PTTS_ATTN_FLASH_MAX_ROWS)
    // measured (profiles/step_bench_flash_r02.txt): 10.7 vs 14.1 us per launch at B = 16, 13.6 vs 15.1 at 64,
    // 20.5 vs 18.9 at 128, equal from 256 on
    const char *fm = getenv("PTTS_ATTN_FLASH_MAX_ROWS");
    const int flash_max_rows = fm ? atoi(fm) : 96;
    if (mode == 2 && rows <= flash_max_rows) {
      mode = 3;
      const size_t need = sizeof(float) * (16 * HD + 32); // the combine area
      if (sm < need)
        sm = need;
    }
  }
  // opt-in (PTTS_ATTN_PAIRS=1): two heads per CTA.  Bit-identical; measured no faster than one head per CTA
  // (54.6 vs 53.2 us per launch at B = 512, 92.8 vs 93.5 at 1024, 35.1 vs 31.3 at 256): what the shared
  // prologue saves, the doubled dependent work per thread gives back.
  if (mode == 2 && (kv.n_heads & 1) == 0 && getenv("PTTS_ATTN_PAIRS") != nullptr) {
    const int score_floats = (max_keys + 15) / 16 * 16;
    dim3 grid2(rows, kv.n_heads / 2);
    const size_t sm2 = sizeof(float) * 2 * (size_t)score_floats;
    if (kv.half)
      launch_k(attention_paged_decode2_kernel<__half>, grid2, dim3(ATT_THREADS), sm2, st, qkv, ldq, out, ldo, row_slot,
               row_pos, cosT, sinT, kv, layer, score_floats, sp);
    else
      launch_k(attention_paged_decode2_kernel<float>, grid2, dim3(ATT_THREADS), sm2, st, qkv, ldq, out, ldo, row_slot,
               row_pos, cosT, sinT, kv, layer, score_floats, sp);
    return 1;
  }
  if (kv.half) {
    if (mode == 3)
      PTTS_DECODE_LAUNCH(__half, 3);
    else if (mode == 2)
      PTTS_DECODE_LAUNCH(__half, 2);
    else if (mode == 1)
      PTTS_DECODE_LAUNCH(__half, 1);
    else
      PTTS_DECODE_LAUNCH(__half, 0);
  } else {
    if (mode == 3)
      PTTS_DECODE_LAUNCH(float, 3);
    else if (mode == 2)
      PTTS_DECODE_LAUNCH(float, 2);
    else if (mode == 1)
      PTTS_DECODE_LAUNCH(float, 1);
    else
      PTTS_DECODE_LAUNCH(float, 0);
  }
#undef PTTS_DECODE_LAUNCH
  return 1;
}

int attention_window(const float *qkv, long ldq, float *out, long ldo, int rows,
                     int D, int n_heads, const int32_t *row_pos, int ctx,
                     cudaStream_t st, SplitOut sp) {
  if (rows <= 0)
    return 0;
  // tcgen05 formulation (attention_tc.cu: S and P.V in TMEM): parity-green, but its synchronous
  // load -> MMA -> softmax -> MMA chain per key block makes it 18-22 % slower than the pipelined
  // register-fragment kernel on B200 (171 vs 140 us for 1024 frames, profiles/README_r02.md), so it is
  // opt-in until it is warp-specialised.
  static const bool use_tc = getenv("PTTS_ATTN_TC") != nullptr;
  if (use_tc && ctx > 0) {
    const int n = attention_window_tc(qkv, ldq, out, ldo, rows, D, n_heads, row_pos, ctx, st, sp);
    if (n > 0)
      return n;
  }
  static const bool no_mma = getenv("PTTS_ATTN_NO_MMA") != nullptr;
  if (!no_mma && ctx > 0) {
    const int n = attention_window_mma(qkv, ldq, out, ldo, rows, D, n_heads, row_pos, ctx, st, sp);
    if (n > 0)
      return n;
  }
  if (ctx > 0 && ctx <= 1024) {
    dim3 grid((rows + MQ - 1) / MQ, n_heads);
    size_t sm = sizeof(float) * (size_t)MQ * (ctx + MQ);
    attention_window_mq_kernel<<<grid, ATT_THREADS, sm, st>>>(qkv, ldq, out, ldo, rows, D, row_pos, ctx, sp);
    return 1;
  }
  dim3 grid(rows, n_heads);
  int mk = ctx > 0 ? ctx : 8192;
  size_t sm = sizeof(float) * (size_t)((mk + 15) / 16 * 16);
  attention_window_kernel<<<grid, ATT_THREADS, sm, st>>>(qkv, ldq, out, ldo, D,
                                                         row_pos, ctx, sp);
  return 1;
}

// ---------------------------------------------------------------------------
// RoPE + KV append
// ---------------------------------------------------------------------------
template <typename KT>
__global__ void __launch_bounds__(256)
    rope_kv_append_kernel(float *qkv, const int32_t *row_slot,
                          const int32_t *row_pos, const float *cosT,
                          const float *sinT, KvGeom kv, int layer) {
  PDL_TRIGGER();
  PDL_WAIT();
  const int r = blockIdx.x;
  const int slot = row_slot[r], pos = row_pos[r];
  if (pos < 0)
    return;
  const int H = kv.n_heads, D = H * HD, half = HD / 2;
  float *row = qkv + (long)r * 3 * D;
  const float *c = cosT + (long)pos * half, *s = sinT + (long)pos * half;
  const int PT = kv.page_tokens;
  const long page_stride = (long)kv.n_layers * 2 * H * PT * HD;
  KT *page = static_cast<KT *>(kv.pool) + (long)kv.pagetab[(long)slot * kv.max_pages + pos / PT] * page_stride;
  const int tok = pos % PT;
  // Every load of a pass is issued before its first store: qkv and the page may alias as far as
  // the compiler knows, so a load-store-load order would cost one dependent memory round trip
  // per element.  Thread p handles pair p of q|k (D pairs in all) and element p of v.
  for (int p0 = 0; p0 < D; p0 += 4 * (int)blockDim.x) {
    float a[4], b[4], cj[4], sj[4], vv[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const int p = p0 + threadIdx.x + k * blockDim.x;
      if (p < D) {
        const int isk = p >= D / 2;
        const int pp = isk ? p - D / 2 : p;
        const int h = pp / half, j = pp % half;
        const float *src = row + (isk ? D : 0) + h * HD + 2 * j;
        a[k] = src[0];
        b[k] = src[1];
        cj[k] = c[j];
        sj[k] = s[j];
        vv[k] = row[2 * D + p];
      }
    }
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const int p = p0 + threadIdx.x + k * blockDim.x;
      if (p < D) {
        const int isk = p >= D / 2;
        const int pp = isk ? p - D / 2 : p;
        const int h = pp / half, j = pp % half;
        const float o0 = rope_even(a[k], b[k], cj[k], sj[k]);
        const float o1 = rope_odd(a[k], b[k], cj[k], sj[k]);
        if (isk) {
          KT *dst = page + ((long)(layer * 2 + 0) * H + h) * PT * HD + (long)tok * HD + 2 * j;
          kv_store(dst, o0);
          kv_store(dst + 1, o1);
        } else {
          row[h * HD + 2 * j] = o0;
          row[h * HD + 2 * j + 1] = o1;
        }
        kv_store(page + ((long)(layer * 2 + 1) * H + p / HD) * PT * HD + (long)tok * HD + p % HD, vv[k]);
      }
    }
  }
}

int rope_kv_append(float *qkv, int rows, const int32_t *row_slot,
                   const int32_t *row_pos, const float *cosT, const float *sinT,
                   const KvGeom &kv, int layer, cudaStream_t st) {
  if (rows <= 0)
    return 0;
  if (kv.half)
    launch_k(rope_kv_append_kernel<__half>, dim3(rows), dim3(256), 0, st, qkv, row_slot, row_pos, cosT, sinT, kv, layer);
  else
    launch_k(rope_kv_append_kernel<float>, dim3(rows), dim3(256), 0, st, qkv, row_slot, row_pos, cosT, sinT, kv, layer);
  return 1;
}

template <typename KT>
__global__ void __launch_bounds__(256)
    kv_write_rows_kernel(const float *rk, const float *rv, int slot, int pos0,
                         KvGeom kv, int layer) {
  const int t = blockIdx.x, pos = pos0 + t;
  const int H = kv.n_heads, D = H * HD, PT = kv.page_tokens;
  const long page_stride = (long)kv.n_layers * 2 * H * PT * HD;
  KT *page = static_cast<KT *>(kv.pool) + (long)kv.pagetab[(long)slot * kv.max_pages + pos / PT] * page_stride;
  const int tok = pos % PT;
  for (int i = threadIdx.x; i < D; i += blockDim.x) {
    const int h = i / HD, d = i % HD;
    kv_store(page + ((long)(layer * 2 + 0) * H + h) * PT * HD + (long)tok * HD + d, rk[(long)t * D + i]);
    kv_store(page + ((long)(layer * 2 + 1) * H + h) * PT * HD + (long)tok * HD + d, rv[(long)t * D + i]);
  }
}

int kv_write_rows(const float *rows_k, const float *rows_v, int T, int slot,
                  int pos0, const KvGeom &kv, int layer, cudaStream_t st) {
  if (T <= 0)
    return 0;
  if (kv.half)
    kv_write_rows_kernel<__half><<<T, 256, 0, st>>>(rows_k, rows_v, slot, pos0, kv, layer);
  else
    kv_write_rows_kernel<float><<<T, 256, 0, st>>>(rows_k, rows_v, slot, pos0, kv, layer);
  return 1;
}

template <typename KT>
__global__ void __launch_bounds__(256)
    kv_read_rows_kernel(float *out, int T, int slot, KvGeom kv, int layer, int which) {
  const int t = blockIdx.x;
  const int H = kv.n_heads, D = H * HD, PT = kv.page_tokens;
  const long page_stride = (long)kv.n_layers * 2 * H * PT * HD;
  const KT *page = static_cast<const KT *>(kv.pool) + (long)kv.pagetab[(long)slot * kv.max_pages + t / PT] * page_stride;
  const int tok = t % PT;
  for (int i = threadIdx.x; i < D; i += blockDim.x) {
    const int h = i / HD, d = i % HD;
    out[((long)h * T + t) * HD + d] = kv_get(page + ((long)(layer * 2 + which) * H + h) * PT * HD + (long)tok * HD + d);
  }
}

int kv_read_rows(float *out, int T, int slot, const KvGeom &kv, int layer,
                 int which, cudaStream_t st) {
  if (T <= 0)
    return 0;
  if (kv.half)
    kv_read_rows_kernel<__half><<<T, 256, 0, st>>>(out, T, slot, kv, layer, which);
  else
    kv_read_rows_kernel<float><<<T, 256, 0, st>>>(out, T, slot, kv, layer, which);
  return 1;
}

// ---- stateful streaming decode: per-request K/V ring of the Mimi transformer ----
__global__ void __launch_bounds__(256)
    kv_ring_append_kernel(const float *qkv, int D, const int32_t *row_apos, const int4 *row_seg, float *ring,
                          long state_stride, long layer_off) {
  const int r = blockIdx.x, p = row_apos[r];
  const int4 sg = row_seg[r];
  if (p < 0 || sg.x < 0)
    return;
  const float4 *src = reinterpret_cast<const float4 *>(qkv + (long)r * 3 * D + D); // K | V of the row, 2 D floats
  float4 *dst = reinterpret_cast<float4 *>(ring + (long)sg.x * state_stride + layer_off + (long)(p % MIMI_RING) * 2 * D);
  for (int i = threadIdx.x; i < 2 * D / 4; i += blockDim.x)
    dst[i] = src[i];
}
int kv_ring_append(const float *qkv, int rows, int D, const int32_t *row_apos, const int4 *row_seg, float *ring,
                   long state_stride, long layer_off, cudaStream_t st) {
  if (rows <= 0)
    return 0;
  kv_ring_append_kernel<<<rows, 256, 0, st>>>(qkv, D, row_apos, row_seg, ring, state_stride, layer_off);
  return 1;
}

// Row r (absolute position p) attends the last min(ctx, p + 1) rows of its request: those at or after the
// segment's first row come from the qkv buffer, the older ones from the request's ring
// (mimi.go:386-441, attention.go:473-484 with the window carried across deliveries).
__global__ void __launch_bounds__(ATT_THREADS)
    attention_window_stream_kernel(const float *qkv, long ldq, float *out, long ldo, int D, const int32_t *row_apos,
                                   const int4 *row_seg, const float *ring, long state_stride, long layer_off, int ctx,
                                   SplitOut sp) {
  extern __shared__ float smem[];
  __shared__ float red[ATT_THREADS / 32];
  __shared__ float4 pv[8 * 16];
  const int r = blockIdx.x, h = blockIdx.y;
  const int pos = row_apos[r];
  float *o = out ? out + (long)r * ldo + h * HD : nullptr;
  if (pos < 0) {
    if (threadIdx.x < 16) {
      const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
      if (o)
        *reinterpret_cast<float4 *>(o + threadIdx.x * 4) = z;
      if (sp.hi)
        split_store4(sp, r, h * HD + threadIdx.x * 4, z);
    }
    return;
  }
  const int4 sg = row_seg[r];
  const int nkeys = (ctx > 0 && pos + 1 > ctx) ? ctx : pos + 1;
  const int first = pos - (nkeys - 1); // absolute position of the oldest allowed key
  const float *rk = ring + (long)(sg.x < 0 ? 0 : sg.x) * state_stride + layer_off + h * HD;
  const float *bk = qkv + D + h * HD;
  auto keyp = [&](int j) {
    const int a = first + j;
    return a >= sg.y ? bk + (long)(sg.z + a - sg.y) * ldq : rk + (long)(a % MIMI_RING) * 2 * D;
  };
  auto valp = [&](int j) {
    const int a = first + j;
    return a >= sg.y ? bk + D + (long)(sg.z + a - sg.y) * ldq : rk + D + (long)(a % MIMI_RING) * 2 * D;
  };
  attend_row(qkv + (long)r * ldq + h * HD, nkeys, keyp, valp, o, smem, red, pv, sp, r, h * HD);
}
int attention_window_stream(const float *qkv, long ldq, float *out, long ldo, int rows, int D, int n_heads,
                            const int32_t *row_apos, const int4 *row_seg, const float *ring, long state_stride,
                            long layer_off, int ctx, cudaStream_t st, SplitOut sp) {
  if (rows <= 0)
    return 0;
  dim3 grid(rows, n_heads);
  const int mk = ctx > 0 ? ctx : 8192;
  const size_t sm = sizeof(float) * (size_t)((mk + 15) / 16 * 16);
  attention_window_stream_kernel<<<grid, ATT_THREADS, sm, st>>>(qkv, ldq, out, ldo, D, row_apos, row_seg, ring,
                                                                state_stride, layer_off, ctx, sp);
  return 1;
}

__global__ void __launch_bounds__(256)
    rope_inplace_kernel(float *qkv, int D, int n_heads, const int32_t *row_pos,
                        const float *cosT, const float *sinT) {
  const int r = blockIdx.x, pos = row_pos[r];
  if (pos < 0)
    return;
  const int half = HD / 2;
  float *row = qkv + (long)r * 3 * D;
  const float *c = cosT + (long)pos * half, *s = sinT + (long)pos * half;
  for (int p = threadIdx.x; p < D; p += blockDim.x) {
    const int isk = p >= D / 2;
    const int pp = isk ? p - D / 2 : p;
    const int h = pp / half, j = pp % half;
    float *src = row + (isk ? D : 0) + h * HD + 2 * j;
    const float a = src[0], b = src[1];
    src[0] = rope_even(a, b, c[j], s[j]);
    src[1] = rope_odd(a, b, c[j], s[j]);
  }
}

int rope_inplace(float *qkv, int rows, int D, int n_heads, const int32_t *row_pos,
                 const float *cosT, const float *sinT, cudaStream_t st) {
  if (rows <= 0)
    return 0;
  rope_inplace_kernel<<<rows, 256, 0, st>>>(qkv, D, n_heads, row_pos, cosT, sinT);
  return 1;
}

} // namespace ptts
