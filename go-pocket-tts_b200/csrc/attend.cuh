// attend.cuh — one (query row, head) of masked attention with 128 threads: scores,
// max / exp / sum, P.V (see attention.cu for the reference semantics).  Shared by the
// stand-alone attention kernels and the fused B=1 frame kernel (step_fused.cu).
#pragma once
#include "kernels.cuh"
#include "split.cuh"
#include <cuda_fp16.h>
#include <math.h>
#include <type_traits>

namespace ptts {

constexpr int HD = 64;       // head_dim (FlowLM 1024/16, Mimi 512/8)
constexpr int ATT_THREADS = 128;

__device__ __forceinline__ float block_reduce_max(float v, float *red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
    v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  if ((threadIdx.x & 31) == 0)
    red[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = red[0];
#pragma unroll
  for (int i = 1; i < ATT_THREADS / 32; i++)
    r = fmaxf(r, red[i]);
  __syncthreads();
  return r;
}
__device__ __forceinline__ float block_reduce_sum(float v, float *red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
    v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0)
    red[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = 0.0f;
#pragma unroll
  for (int i = 0; i < ATT_THREADS / 32; i++)
    r += red[i];
  __syncthreads();
  return r;
}

// RoPE rotation of one interleaved pair (rope.go:81-105: x0 c - x1 s, x0 s + x1 c, every product and the sum
// rounded separately — the reference's Go has no FMA contraction on amd64).  Intrinsics, so that every kernel that
// rotates (prefill append, decode attention variants, the batch-1 frame kernels) produces the same bits whatever
// the surrounding code looks like to the compiler.
__device__ __forceinline__ float rope_even(float x0, float x1, float c, float s) {
  return __fsub_rn(__fmul_rn(x0, c), __fmul_rn(x1, s));
}
__device__ __forceinline__ float rope_odd(float x0, float x1, float c, float s) {
  return __fadd_rn(__fmul_rn(x0, s), __fmul_rn(x1, c));
}

// Element type of a K/V row: fp32 (Mimi: rows of the qkv GEMM output; FlowLM cache in exact mode)
// or fp16 (the FlowLM paged cache by default: the cache stream is what bounds the batched AR step,
// and 11 mantissa bits on K and V keep the attention output ~1e-4 relative, inside the 1e-3 latent
// tolerance; the arithmetic — scores, softmax, P.V — stays fp32).  Lane `part` of a key's 8 lanes
// owns elements [8 part, 8 part + 8); a value lane owns 4 consecutive elements.
__device__ __forceinline__ void kv_load8(const float *p, float (&v)[8]) {
  const float4 a = *reinterpret_cast<const float4 *>(p), b = *reinterpret_cast<const float4 *>(p + 4);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void kv_load8(const __half *p, float (&v)[8]) {
  const uint4 u = *reinterpret_cast<const uint4 *>(p);
  const float2 a = __half22float2(*reinterpret_cast<const __half2 *>(&u.x));
  const float2 b = __half22float2(*reinterpret_cast<const __half2 *>(&u.y));
  const float2 c = __half22float2(*reinterpret_cast<const __half2 *>(&u.z));
  const float2 d = __half22float2(*reinterpret_cast<const __half2 *>(&u.w));
  v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y; v[4] = c.x; v[5] = c.y; v[6] = d.x; v[7] = d.y;
}
__device__ __forceinline__ float4 kv_load4(const float *p) { return *reinterpret_cast<const float4 *>(p); }
__device__ __forceinline__ float4 kv_load4(const __half *p) {
  const uint2 u = *reinterpret_cast<const uint2 *>(p);
  const float2 a = __half22float2(*reinterpret_cast<const __half2 *>(&u.x));
  const float2 b = __half22float2(*reinterpret_cast<const __half2 *>(&u.y));
  return make_float4(a.x, a.y, b.x, b.y);
}
__device__ __forceinline__ void kv_store(float *p, float v) { *p = v; }
__device__ __forceinline__ void kv_store(__half *p, float v) { *p = __float2half_rn(v); }
__device__ __forceinline__ float kv_get(const float *p) { return *p; }
__device__ __forceinline__ float kv_get(const __half *p) { return __half2float(*p); }

// KeyFn(j) -> pointer to the 64 elements of key j; ValFn(j) likewise for values.
template <class KeyFn, class ValFn>
__device__ __forceinline__ void attend_row(const float *q, int nkeys, KeyFn keyp,
                                           ValFn valp, float *out, float *scores,
                                           float *red, float4 *pv, const SplitOut &sp, long sp_row,
                                           int sp_col, uint2 *ll_out = nullptr, unsigned ll_tag = 0) {
  const int tid = threadIdx.x;
  const float scale = 0.125f; // 1/sqrt(64)
  // phase 1
  const int part = tid & 7, kslot = tid >> 3; // 16 keys per pass
  const float4 q0 = *reinterpret_cast<const float4 *>(q + part * 8);
  const float4 q1 = *reinterpret_cast<const float4 *>(q + part * 8 + 4);
  float lmax = -INFINITY;
  for (int j0 = 0; j0 < nkeys; j0 += 16) {
    const int j = j0 + kslot;
    float s = 0.0f;
    if (j < nkeys) {
      float k[8];
      kv_load8(keyp(j) + part * 8, k);
      s = q0.x * k[0];
      s = fmaf(q0.y, k[1], s);
      s = fmaf(q0.z, k[2], s);
      s = fmaf(q0.w, k[3], s);
      s = fmaf(q1.x, k[4], s);
      s = fmaf(q1.y, k[5], s);
      s = fmaf(q1.z, k[6], s);
      s = fmaf(q1.w, k[7], s);
    }
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    s += __shfl_xor_sync(0xffffffffu, s, 4);
    if (j < nkeys) {
      s *= scale;
      if (part == 0)
        scores[j] = s;
      lmax = fmaxf(lmax, s);
    }
  }
  const float gmax = block_reduce_max(lmax, red); // includes __syncthreads
  // phase 2
  float lsum = 0.0f;
  for (int j = tid; j < nkeys; j += ATT_THREADS) {
    float e = expf(scores[j] - gmax);
    scores[j] = e;
    lsum += e;
  }
  const float gsum = block_reduce_sum(lsum, red);
  const float inv = 1.0f / gsum;
  // phase 3
  const int c4 = tid & 15, grp = tid >> 4; // 8 key groups
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  // four value rows (and their page-table lookups) in flight per thread: the loop is ~nkeys/8 dependent
  // round trips otherwise.  Same summation order as the plain loop.
  int j = grp;
  for (; j + 24 < nkeys; j += 32) {
    const float4 v0 = kv_load4(valp(j) + c4 * 4), v1 = kv_load4(valp(j + 8) + c4 * 4);
    const float4 v2 = kv_load4(valp(j + 16) + c4 * 4), v3 = kv_load4(valp(j + 24) + c4 * 4);
    const float w0 = scores[j] * inv, w1 = scores[j + 8] * inv, w2 = scores[j + 16] * inv, w3 = scores[j + 24] * inv;
    acc.x = fmaf(w0, v0.x, acc.x); acc.y = fmaf(w0, v0.y, acc.y); acc.z = fmaf(w0, v0.z, acc.z); acc.w = fmaf(w0, v0.w, acc.w);
    acc.x = fmaf(w1, v1.x, acc.x); acc.y = fmaf(w1, v1.y, acc.y); acc.z = fmaf(w1, v1.z, acc.z); acc.w = fmaf(w1, v1.w, acc.w);
    acc.x = fmaf(w2, v2.x, acc.x); acc.y = fmaf(w2, v2.y, acc.y); acc.z = fmaf(w2, v2.z, acc.z); acc.w = fmaf(w2, v2.w, acc.w);
    acc.x = fmaf(w3, v3.x, acc.x); acc.y = fmaf(w3, v3.y, acc.y); acc.z = fmaf(w3, v3.z, acc.z); acc.w = fmaf(w3, v3.w, acc.w);
  }
  for (; j < nkeys; j += 8) {
    const float w = scores[j] * inv;
    const float4 v = kv_load4(valp(j) + c4 * 4);
    acc.x = fmaf(w, v.x, acc.x);
    acc.y = fmaf(w, v.y, acc.y);
    acc.z = fmaf(w, v.z, acc.z);
    acc.w = fmaf(w, v.w, acc.w);
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
    if (ll_out) { // (value, tag) words for the flag-in-data exchange of step_fused.cu
      const float rv[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
      for (int e = 0; e < 4; e++)
        asm volatile("st.relaxed.gpu.global.v2.u32 [%0], {%1, %2};" ::"l"(ll_out + tid * 4 + e),
                     "r"(__float_as_uint(rv[e])), "r"(ll_tag)
                     : "memory");
    }
  }
}

// One (row, head) of decode attention in ONE pass over 16-token pages (flash-decoding inside the CTA): the 16
// groups of 8 lanes are 16 independent softmax streams — warp w takes pages w, w + 4, ..., group g of the warp
// rows g, g + 4, g + 8, g + 12 of the page — each with its own running max, denominator and 64-wide accumulator
// (8 elements per lane); K and V rows of a page are requested together, and the streams are combined once at the
// end.  One block barrier instead of six and half the dependent memory round trips per CTA: this is what bounds
// the kernel when there are too few (row, head) pairs to hide a CTA's latency chain (small batches, and the
// batch-1 frame kernels, where 16 CTAs run this phase while the rest of the GPU waits).  The
// summation order differs from attend_row's (not bit-identical to it; same tolerance against the oracle).
// 8 consecutive K / V elements as loaded (registers: 4 for fp16, 8 for fp32), unpacked to fp32 where they are used
template <typename KT>
struct KvRaw8;
template <>
struct KvRaw8<__half> {
  uint4 u;
  __device__ __forceinline__ void load(const __half *p) { u = *reinterpret_cast<const uint4 *>(p); }
  __device__ __forceinline__ void unpack(float (&v)[8]) const {
    const float2 a = __half22float2(*reinterpret_cast<const __half2 *>(&u.x));
    const float2 b = __half22float2(*reinterpret_cast<const __half2 *>(&u.y));
    const float2 c = __half22float2(*reinterpret_cast<const __half2 *>(&u.z));
    const float2 d = __half22float2(*reinterpret_cast<const __half2 *>(&u.w));
    v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y; v[4] = c.x; v[5] = c.y; v[6] = d.x; v[7] = d.y;
  }
};
template <>
struct KvRaw8<float> {
  float4 a, b;
  __device__ __forceinline__ void load(const float *p) {
    a = *reinterpret_cast<const float4 *>(p);
    b = *reinterpret_cast<const float4 *>(p + 4);
  }
  __device__ __forceinline__ void unpack(float (&v)[8]) const {
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  }
};

template <typename KT>
__device__ __forceinline__ void attend_row_flash16(const float *q, int nkeys, const KT *kbase, const KT *vbase,
                                                   const int32_t *s_pt, long page_stride, float *out, float *comb,
                                                   const SplitOut &sp, long sp_row, int sp_col,
                                                   uint2 *ll_out = nullptr, unsigned ll_tag = 0) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int part = lane & 7, g = lane >> 3;
  const float scale = 0.125f; // 1/sqrt(64)
  const int npages = (nkeys + 15) >> 4;
  float qv[8];
  {
    const float4 a = *reinterpret_cast<const float4 *>(q + part * 8), b = *reinterpret_cast<const float4 *>(q + part * 8 + 4);
    qv[0] = a.x; qv[1] = a.y; qv[2] = a.z; qv[3] = a.w; qv[4] = b.x; qv[5] = b.y; qv[6] = b.z; qv[7] = b.w;
  }
  float m = -INFINITY, l = 0.0f, acc[8];
#pragma unroll
  for (int e = 0; e < 8; e++)
    acc[e] = 0.0f;
  const long roff = (long)g * HD + part * 8;
  for (int p = warp; p < npages; p += 4) {
    const long base = (long)s_pt[p] * page_stride + roff;
    const int left = nkeys - p * 16 - g; // rows g + 4 i of this page are valid while 4 i < left
    KvRaw8<KT> kr[4], vr[4];
#pragma unroll
    for (int i = 0; i < 4; i++)
      if (4 * i < left) {
        kr[i].load(kbase + base + (long)(4 * i) * HD);
        vr[i].load(vbase + base + (long)(4 * i) * HD);
      }
    float sc[4], pm = -INFINITY;
#pragma unroll
    for (int i = 0; i < 4; i++) {
      float s = 0.0f;
      if (4 * i < left) {
        float kf[8];
        kr[i].unpack(kf);
#pragma unroll
        for (int e = 0; e < 8; e++)
          s = fmaf(qv[e], kf[e], s);
      }
      s += __shfl_xor_sync(0xffffffffu, s, 1);
      s += __shfl_xor_sync(0xffffffffu, s, 2);
      s += __shfl_xor_sync(0xffffffffu, s, 4);
      sc[i] = 4 * i < left ? s * scale : -INFINITY;
      pm = fmaxf(pm, sc[i]);
    }
    if (left > 0) { // (a stream's page always has its first row unless the page ends before row g)
      const float mn = fmaxf(m, pm);
      const float rescale = expf(m - mn); // 0 on the stream's first page (m = -inf)
      l *= rescale;
#pragma unroll
      for (int e = 0; e < 8; e++)
        acc[e] *= rescale;
#pragma unroll
      for (int i = 0; i < 4; i++)
        if (4 * i < left) {
          const float pe = expf(sc[i] - mn);
          l += pe;
          float vf[8];
          vr[i].unpack(vf);
#pragma unroll
          for (int e = 0; e < 8; e++)
            acc[e] = fmaf(pe, vf[e], acc[e]);
        }
      m = mn;
    }
  }
  // ---- combine the 16 streams: comb = [16][64] accumulators | [16] max | [16] denominators ----
  const int stream = warp * 4 + g;
  float *cm = comb + 16 * HD, *cl = cm + 16;
#pragma unroll
  for (int e = 0; e < 8; e += 4)
    *reinterpret_cast<float4 *>(comb + stream * HD + part * 8 + e) = make_float4(acc[e], acc[e + 1], acc[e + 2], acc[e + 3]);
  if (part == 0) {
    cm[stream] = m;
    cl[stream] = l;
  }
  __syncthreads();
  if (tid < 16) {
    float M = -INFINITY;
#pragma unroll
    for (int s2 = 0; s2 < 16; s2++)
      M = fmaxf(M, cm[s2]);
    float den = 0.0f;
    float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int s2 = 0; s2 < 16; s2++) {
      const float ms = cm[s2];
      if (ms == -INFINITY)
        continue; // a stream that saw no key
      const float wgt = expf(ms - M);
      den = fmaf(wgt, cl[s2], den);
      const float4 a = *reinterpret_cast<const float4 *>(comb + s2 * HD + tid * 4);
      r.x = fmaf(wgt, a.x, r.x);
      r.y = fmaf(wgt, a.y, r.y);
      r.z = fmaf(wgt, a.z, r.z);
      r.w = fmaf(wgt, a.w, r.w);
    }
    const float inv = 1.0f / den;
    r.x *= inv; r.y *= inv; r.z *= inv; r.w *= inv;
    if (out)
      *reinterpret_cast<float4 *>(out + tid * 4) = r;
    if (sp.hi)
      split_store4(sp, sp_row, sp_col + tid * 4, r);
    if (ll_out) { // (value, tag) words for the flag-in-data exchange of step_fused.cu
      const float rv[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
      for (int e = 0; e < 4; e++)
        asm volatile("st.relaxed.gpu.global.v2.u32 [%0], {%1, %2};" ::"l"(ll_out + tid * 4 + e),
                     "r"(__float_as_uint(rv[e])), "r"(ll_tag)
                     : "memory");
    }
  }
}


} // namespace ptts
