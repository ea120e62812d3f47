// attention_mma.cu — sliding-window causal attention of the Mimi decoder transformer on
// tensor cores (reference: internal/native/mimi.go:386-441 over
// internal/runtime/ops/attention.go:307-484, context 250, head_dim 64).
//
// Flash-style: a CTA (4 warps) owns 64 consecutive query rows of one head and walks the
// key rows of their windows in blocks of 64.  Per block and warp (16 queries):
//   S = Q K^T  (16 x 64 x 64)  ->  masked online softmax  ->  O += P V  (16 x 64 x 64)
// Q, K, V and P are split x = hi + lo (bf16 each) and every product is three
// mma.sync.m16n8k16 bf16 instructions (hi*hi + hi*lo + lo*hi, fp32 accumulate), so the
// result keeps ~16 mantissa bits like the tcgen05 GEMM path: plain bf16 scores would not
// hold the attention tolerance (2e-4, ops/tolerance.go:13-24).
//
// This is the register-fragment (legacy) tensor path, chosen because Q/K/V arrive as fp32
// rows of the qkv GEMM output and P lives in registers; the tcgen05/TMEM formulation
// (S and O in TMEM, V^T by TMA) is the follow-up noted in DESIGN.md §6.
#include "kernels.cuh"
#include "split.cuh"
#include <math.h>

namespace ptts {

// exp as ex2.approx.ftz(x * log2 e): __expf is the same product and MUFU wrapped in denormal scaling (two more
// FMUL and an FSETP per call); the softmax here issues one exp per (query, key) pair.  exp(-inf) = 0 as before.
__device__ __forceinline__ float exp_ftz(float x) {
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(__fmul_rn(x, 1.4426950216293335f)));
  return e;
}

namespace {

constexpr int AM_HD = 64;      // head_dim
constexpr int AM_TQ = 64;      // queries per CTA (16 per warp)
constexpr int AM_TK = 64;      // keys per block
constexpr int AM_LD = AM_HD + 8; // smem row stride in bf16 (144 B: conflict-free ldmatrix)
constexpr int AM_THREADS = 128;

__device__ __forceinline__ void mma_bf16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_trans(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}

__global__ void __launch_bounds__(AM_THREADS)
    attention_window_mma_kernel(const float *qkv, long ldq, float *out, long ldo, int rows, int D,
                                const int32_t *row_pos, int ctx, SplitOut sp) {
  __shared__ __align__(16) uint16_t sK[2][AM_TK][AM_LD]; // [hi|lo][key][dim]
  __shared__ __align__(16) uint16_t sV[2][AM_TK][AM_LD];
  __shared__ int s_lo[AM_TQ], s_n[AM_TQ];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  const int h = blockIdx.y;
  const int r0 = blockIdx.x * AM_TQ;
  // allowed keys of query r (local position p): the last n = min(ctx, p+1) rows
  // (0 <= pos_q - pos_k < ctx inside the segment, attention.go:473-484)
  if (tid < AM_TQ) {
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
  __syncthreads();
  int lo_min = 0x7fffffff, hi_max = -1;
  for (int i = 0; i < AM_TQ; i++)
    if (s_n[i] > 0) {
      lo_min = min(lo_min, s_lo[i]);
      hi_max = r0 + i;
    }
  // this warp's 16 queries; rows g and g+8 of the fragment
  const int qa = r0 + warp * 16 + g, qb = qa + 8;
  const int na = s_n[warp * 16 + g], nb = s_n[warp * 16 + g + 8];
  const int loa = s_lo[warp * 16 + g], lob = s_lo[warp * 16 + g + 8];
  int wlo = 0x7fffffff, whi = -1; // key range any query of this warp needs
  for (int i = 0; i < 16; i++)
    if (s_n[warp * 16 + i] > 0) {
      wlo = min(wlo, s_lo[warp * 16 + i]);
      whi = r0 + warp * 16 + i;
    }

  // Q fragments (scaled by 1/sqrt(64) = 2^-3, exact), split once
  uint32_t qh[4][4], ql[4][4];
  {
    const float *qpa = qkv + (long)min(qa, rows - 1) * ldq + h * AM_HD;
    const float *qpb = qkv + (long)min(qb, rows - 1) * ldq + h * AM_HD;
#pragma unroll
    for (int ks = 0; ks < 4; ks++) {
      const float2 a0 = *reinterpret_cast<const float2 *>(qpa + ks * 16 + 2 * t);
      const float2 a1 = *reinterpret_cast<const float2 *>(qpb + ks * 16 + 2 * t);
      const float2 a2 = *reinterpret_cast<const float2 *>(qpa + ks * 16 + 8 + 2 * t);
      const float2 a3 = *reinterpret_cast<const float2 *>(qpb + ks * 16 + 8 + 2 * t);
      split2(a0.x * 0.125f, a0.y * 0.125f, qh[ks][0], ql[ks][0]);
      split2(a1.x * 0.125f, a1.y * 0.125f, qh[ks][1], ql[ks][1]);
      split2(a2.x * 0.125f, a2.y * 0.125f, qh[ks][2], ql[ks][2]);
      split2(a3.x * 0.125f, a3.y * 0.125f, qh[ks][3], ql[ks][3]);
    }
  }
  float o[8][4];
#pragma unroll
  for (int i = 0; i < 8; i++)
#pragma unroll
    for (int j = 0; j < 4; j++)
      o[i][j] = 0.0f;
  float ma = -INFINITY, mb = -INFINITY, la = 0.0f, lb = 0.0f;

  const float *kbase = qkv + D + h * AM_HD, *vbase = qkv + 2 * D + h * AM_HD;
  const uint32_t sk_hi = (uint32_t)__cvta_generic_to_shared(&sK[0][0][0]);
  const uint32_t sk_lo = (uint32_t)__cvta_generic_to_shared(&sK[1][0][0]);
  const uint32_t sv_hi = (uint32_t)__cvta_generic_to_shared(&sV[0][0][0]);
  const uint32_t sv_lo = (uint32_t)__cvta_generic_to_shared(&sV[1][0][0]);
  // per-lane ldmatrix offsets (bytes)
  // K (non-transposed, B of S = Q K^T): matrix id = lane/8 -> dims (ks*16 + id*8), key row = lane%8
  const uint32_t k_lane_off = (uint32_t)(((lane & 7) * AM_LD + (lane >> 3) * 8) * 2);
  // V (transposed, B of O = P V): id&1 -> keys +8, id>>1 -> dims +8
  const uint32_t v_lane_off = (uint32_t)(((((lane >> 3) & 1) * 8 + (lane & 7)) * AM_LD + (lane >> 4) * 8) * 2);

  if (hi_max >= 0) {
    for (int jb = lo_min; jb <= hi_max; jb += AM_TK) {
      __syncthreads(); // previous block's fragments are consumed
      // ---- K, V block: fp32 rows -> split bf16 smem (rows outside [0, rows) are zero) ----
#pragma unroll
      for (int it = 0; it < (AM_TK * AM_HD / 4) / AM_THREADS; it++) {
        const int idx = it * AM_THREADS + tid;
        const int key = idx >> 4, c4 = (idx & 15) * 4;
        const int j = jb + key;
        float4 kv4 = make_float4(0.f, 0.f, 0.f, 0.f), vv4 = kv4;
        if (j >= 0 && j < rows && j <= hi_max) {
          kv4 = *reinterpret_cast<const float4 *>(kbase + (long)j * ldq + c4);
          vv4 = *reinterpret_cast<const float4 *>(vbase + (long)j * ldq + c4);
        }
        uint2 hi, lo;
        split2(kv4.x, kv4.y, hi.x, lo.x);
        split2(kv4.z, kv4.w, hi.y, lo.y);
        *reinterpret_cast<uint2 *>(&sK[0][key][c4]) = hi;
        *reinterpret_cast<uint2 *>(&sK[1][key][c4]) = lo;
        split2(vv4.x, vv4.y, hi.x, lo.x);
        split2(vv4.z, vv4.w, hi.y, lo.y);
        *reinterpret_cast<uint2 *>(&sV[0][key][c4]) = hi;
        *reinterpret_cast<uint2 *>(&sV[1][key][c4]) = lo;
      }
      __syncthreads();
      if (whi < jb || wlo > jb + AM_TK - 1)
        continue; // no query of this warp sees this block (warp-uniform)
      // ---- S = Q K^T ----
      float s[8][4];
#pragma unroll
      for (int nt = 0; nt < 8; nt++) {
        s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.0f;
#pragma unroll
        for (int kp = 0; kp < 2; kp++) { // two k-steps per ldmatrix.x4
          uint32_t bh[4], bl[4];
          const uint32_t off = (uint32_t)((nt * 8 * AM_LD + kp * 32) * 2) + k_lane_off;
          ldsm_x4(bh, sk_hi + off);
          ldsm_x4(bl, sk_lo + off);
          mma_bf16(s[nt], ql[2 * kp], bh[0], bh[1]);
          mma_bf16(s[nt], qh[2 * kp], bl[0], bl[1]);
          mma_bf16(s[nt], qh[2 * kp], bh[0], bh[1]);
          mma_bf16(s[nt], ql[2 * kp + 1], bh[2], bh[3]);
          mma_bf16(s[nt], qh[2 * kp + 1], bl[2], bl[3]);
          mma_bf16(s[nt], qh[2 * kp + 1], bh[2], bh[3]);
        }
      }
      // ---- mask + online softmax (rows qa: s[..][0,1], qb: s[..][2,3]) ----
      float bma = -INFINITY, bmb = -INFINITY;
#pragma unroll
      for (int nt = 0; nt < 8; nt++) {
#pragma unroll
        for (int c = 0; c < 2; c++) {
          const int j = jb + nt * 8 + 2 * t + c;
          const bool oka = na > 0 && j >= loa && j <= qa;
          const bool okb = nb > 0 && j >= lob && j <= qb;
          s[nt][c] = oka ? s[nt][c] : -INFINITY;
          s[nt][2 + c] = okb ? s[nt][2 + c] : -INFINITY;
          bma = fmaxf(bma, s[nt][c]);
          bmb = fmaxf(bmb, s[nt][2 + c]);
        }
      }
      bma = fmaxf(bma, __shfl_xor_sync(0xffffffffu, bma, 1));
      bma = fmaxf(bma, __shfl_xor_sync(0xffffffffu, bma, 2));
      bmb = fmaxf(bmb, __shfl_xor_sync(0xffffffffu, bmb, 1));
      bmb = fmaxf(bmb, __shfl_xor_sync(0xffffffffu, bmb, 2));
      const float mna = fmaxf(ma, bma), mnb = fmaxf(mb, bmb);
      // a row that has seen no allowed key yet keeps m = -inf: all its p are 0
      const float alpha_a = mna == -INFINITY ? 1.0f : exp_ftz(ma - mna);
      const float alpha_b = mnb == -INFINITY ? 1.0f : exp_ftz(mb - mnb);
      const float sha = mna == -INFINITY ? 0.0f : mna, shb = mnb == -INFINITY ? 0.0f : mnb;
      ma = mna;
      mb = mnb;
      la *= alpha_a;
      lb *= alpha_b;
#pragma unroll
      for (int nd = 0; nd < 8; nd++) {
        o[nd][0] *= alpha_a; o[nd][1] *= alpha_a;
        o[nd][2] *= alpha_b; o[nd][3] *= alpha_b;
      }
#pragma unroll
      for (int nt = 0; nt < 8; nt++) {
        s[nt][0] = exp_ftz(s[nt][0] - sha); // exp(-inf) = 0 for masked keys
        s[nt][1] = exp_ftz(s[nt][1] - sha);
        s[nt][2] = exp_ftz(s[nt][2] - shb);
        s[nt][3] = exp_ftz(s[nt][3] - shb);
        la += s[nt][0] + s[nt][1];
        lb += s[nt][2] + s[nt][3];
      }
      // ---- O += P V ----
#pragma unroll
      for (int kk = 0; kk < 4; kk++) { // 16 keys per k-step = S tiles 2kk, 2kk+1
        uint32_t ph[4], pl[4];
        split2(s[2 * kk][0], s[2 * kk][1], ph[0], pl[0]);
        split2(s[2 * kk][2], s[2 * kk][3], ph[1], pl[1]);
        split2(s[2 * kk + 1][0], s[2 * kk + 1][1], ph[2], pl[2]);
        split2(s[2 * kk + 1][2], s[2 * kk + 1][3], ph[3], pl[3]);
#pragma unroll
        for (int np = 0; np < 4; np++) { // two 8-dim output tiles per ldmatrix.x4.trans
          uint32_t vh[4], vl[4];
          const uint32_t off = (uint32_t)((kk * 16 * AM_LD + np * 16) * 2) + v_lane_off;
          ldsm_x4_trans(vh, sv_hi + off);
          ldsm_x4_trans(vl, sv_lo + off);
          mma_bf16(o[2 * np], pl, vh[0], vh[1]);
          mma_bf16(o[2 * np], ph, vl[0], vl[1]);
          mma_bf16(o[2 * np], ph, vh[0], vh[1]);
          mma_bf16(o[2 * np + 1], pl, vh[2], vh[3]);
          mma_bf16(o[2 * np + 1], ph, vl[2], vl[3]);
          mma_bf16(o[2 * np + 1], ph, vh[2], vh[3]);
        }
      }
    }
  }
  // ---- normalise and store (an all-masked row is zeros) ----
  la += __shfl_xor_sync(0xffffffffu, la, 1);
  la += __shfl_xor_sync(0xffffffffu, la, 2);
  lb += __shfl_xor_sync(0xffffffffu, lb, 1);
  lb += __shfl_xor_sync(0xffffffffu, lb, 2);
  const float ia = la > 0.0f ? 1.0f / la : 0.0f, ib = lb > 0.0f ? 1.0f / lb : 0.0f;
#pragma unroll
  for (int nd = 0; nd < 8; nd++) {
    const int col = h * AM_HD + nd * 8 + 2 * t;
    const float2 va = make_float2(o[nd][0] * ia, o[nd][1] * ia);
    const float2 vb = make_float2(o[nd][2] * ib, o[nd][3] * ib);
    if (qa < rows) {
      if (out)
        *reinterpret_cast<float2 *>(out + (long)qa * ldo + col) = va;
      if (sp.hi) {
        uint32_t hi, lo;
        split2(va.x, va.y, hi, lo);
        uint16_t *p = sp.hi + (long)qa * sp.ld + col;
        *reinterpret_cast<uint32_t *>(p) = hi;
        *reinterpret_cast<uint32_t *>(p + sp.lo_off) = lo;
      }
    }
    if (qb < rows) {
      if (out)
        *reinterpret_cast<float2 *>(out + (long)qb * ldo + col) = vb;
      if (sp.hi) {
        uint32_t hi, lo;
        split2(vb.x, vb.y, hi, lo);
        uint16_t *p = sp.hi + (long)qb * sp.ld + col;
        *reinterpret_cast<uint32_t *>(p) = hi;
        *reinterpret_cast<uint32_t *>(p + sp.lo_off) = lo;
      }
    }
  }
}

} // namespace

// head_dim 64, ctx > 0.  Returns 0 (nothing launched) when the shape is not supported so
// the caller can take the CUDA-core kernel.
int attention_window_mma(const float *qkv, long ldq, float *out, long ldo, int rows, int D, int n_heads,
                         const int32_t *row_pos, int ctx, cudaStream_t st, SplitOut sp) {
  if (rows <= 0 || ctx <= 0 || D != n_heads * AM_HD || (ldq & 3) != 0)
    return 0;
  dim3 grid((rows + AM_TQ - 1) / AM_TQ, n_heads);
  attention_window_mma_kernel<<<grid, AM_THREADS, 0, st>>>(qkv, ldq, out, ldo, rows, D, row_pos, ctx, sp);
  return 1;
}

} // namespace ptts
