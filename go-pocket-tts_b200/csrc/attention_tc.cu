// attention_tc.cu — sliding-window causal attention of the Mimi decoder transformer on the 5th-gen
// tensor cores: tcgen05.mma with S and P.V accumulators in TMEM (reference: internal/native/
// mimi.go:386-441 over internal/runtime/ops/attention.go:307-484, context 250, head_dim 64).
//
// Flash-style.  A CTA (4 warps) owns 128 consecutive query rows of one head; thread t is query row
// t and TMEM lane t.  It walks the key rows of the tile's windows in blocks of 64:
//   K block, V block: fp32 rows of the qkv GEMM output -> split bf16 (hi, lo) -> shared memory as
//       128-byte-swizzled [key][dim] rows: K is the K-major B operand of S = Q K^T, V the MN-major
//       B operand of P.V (no transpose on the way in);
//   S = Q K^T   : tcgen05.mma M=128 N=64, 4 k-steps x 3 split products, accumulator TMEM cols 0..63;
//   softmax     : thread t reads ITS row of S with tcgen05.ld (64 values), masks (0 <= pos_q - pos_k
//       < ctx inside the segment, attention.go:473-484), online max / sum in registers, writes its
//       row of P (split bf16) into shared memory as the next A operand;
//   O += P V    : tcgen05.mma M=128 N=64 (dims), K = 64 keys, accumulator TMEM cols 64..127, read
//       back and added to the thread's running O row (rescaled by the online-softmax factor).
// Every product is three bf16 MMAs (hi*hi + hi*lo + lo*hi, fp32 accumulate), so scores and outputs
// keep ~16 mantissa bits like the GEMM path.  96 KB of shared memory and 128 TMEM columns per CTA:
// two CTAs per SM overlap each other's load -> MMA -> softmax chain.
#include "kernels.cuh"
#include "split.cuh"
#include "tc_common.cuh"
#include <math.h>

namespace ptts {

namespace {

constexpr int AT_Q = 128;  // queries per CTA
constexpr int AT_K = 64;   // keys per block
constexpr int AT_HD = 64;  // head_dim
constexpr int AT_THREADS = 128;
constexpr int AT_QT = AT_Q * 128;  // bytes of a [128 x 64] bf16 tile
constexpr int AT_KT = AT_K * 128;  // bytes of a [64 x 64] bf16 tile
constexpr int AT_SMEM = 2 * AT_QT + 4 * AT_KT + 2 * AT_QT + 1024 + 64;

// byte offset of 16-byte chunk c8 (8 bf16) of row `row` in a K-major 128B-swizzled tile
__device__ __forceinline__ uint32_t swz(int row, int c8) {
  return (uint32_t)((row >> 3) * 1024 + (row & 7) * 128 + ((c8 ^ (row & 7)) << 4));
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,"
      "%28,%29,%30,%31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// D[128 x 64] (+)= A[128 x 64] B with split operands: 4 k-steps x 3 products.  A is K-major.  B is
// K-major ([n][k] rows: S = Q K^T, the K block as loaded) or, B_MN, MN-major ([k][n] rows: P V with the
// V block as loaded — no transpose in shared memory): a k-step of 16 then advances the descriptor by
// 16 rows (2048 B, two 8-row swizzle atoms) instead of 32 bytes inside the swizzled row.
template <bool B_MN>
__device__ __forceinline__ void mma_split(uint32_t tmem_d, uint32_t a_hi, uint32_t a_lo, uint32_t b_hi, uint32_t b_lo) {
  constexpr uint32_t idesc = make_idesc(64) | (B_MN ? (1u << 16) : 0u); // bit 16: B is MN-major
#pragma unroll
  for (int k = 0; k < 4; k++) {
    const uint32_t ko = k * 32, kb = B_MN ? k * 2048 : k * 32;
    const uint64_t dah = make_smem_desc(a_hi + ko), dal = make_smem_desc(a_lo + ko);
    const uint64_t dbh = make_smem_desc(b_hi + kb), dbl = make_smem_desc(b_lo + kb);
    umma_bf16(tmem_d, dal, dbh, idesc, k != 0);
    umma_bf16(tmem_d, dah, dbl, idesc, 1);
    umma_bf16(tmem_d, dah, dbh, idesc, 1);
  }
}

__global__ void __launch_bounds__(AT_THREADS)
    attention_window_tc_kernel(const float *__restrict__ qkv, long ldq, float *__restrict__ out, long ldo, int rows, int D,
                               const int32_t *__restrict__ row_pos, int ctx, SplitOut sp) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t *sQh = smem, *sQl = sQh + AT_QT;
  uint8_t *sKh = sQl + AT_QT, *sKl = sKh + AT_KT, *sVh = sKl + AT_KT, *sVl = sVh + AT_KT;
  uint8_t *sPh = sVl + AT_KT, *sPl = sPh + AT_QT;
  uint64_t *bar_s = reinterpret_cast<uint64_t *>(sPl + AT_QT);
  uint64_t *bar_o = bar_s + 1;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bar_o + 1);
  __shared__ int s_lo_min, s_hi_max;

  const int tid = threadIdx.x, warp = tid >> 5;
  const int h = blockIdx.y, r0 = blockIdx.x * AT_Q;
  const int r = r0 + tid; // this thread's query row
  // allowed keys of query r (local position p): the last n = min(ctx, p + 1) rows of its segment
  int n = 0, lo = r;
  if (r < rows) {
    const int pos = row_pos[r];
    if (pos >= 0) {
      n = (ctx > 0 && pos + 1 > ctx) ? ctx : pos + 1;
      lo = r - (n - 1);
    }
  }
  if (tid == 0) {
    s_lo_min = 0x7fffffff;
    s_hi_max = -1;
    mbar_init(bar_s, 1);
    mbar_init(bar_o, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(128)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  __syncthreads();
  if (n > 0) {
    atomicMin(&s_lo_min, lo);
    atomicMax(&s_hi_max, r);
  }
  // Q row (scaled by 1/sqrt(64) = 2^-3, exact), split once into the swizzled A tile
  {
    const float *qp = qkv + (long)min(r, rows - 1) * ldq + h * AT_HD;
#pragma unroll
    for (int c8 = 0; c8 < 8; c8++) {
      float4 u = make_float4(0.f, 0.f, 0.f, 0.f), v = u;
      if (r < rows) {
        u = *reinterpret_cast<const float4 *>(qp + c8 * 8);
        v = *reinterpret_cast<const float4 *>(qp + c8 * 8 + 4);
      }
      u.x *= 0.125f; u.y *= 0.125f; u.z *= 0.125f; u.w *= 0.125f;
      v.x *= 0.125f; v.y *= 0.125f; v.z *= 0.125f; v.w *= 0.125f;
      uint4 hi, lo4;
      split8(u, v, hi, lo4);
      *reinterpret_cast<uint4 *>(sQh + swz(tid, c8)) = hi;
      *reinterpret_cast<uint4 *>(sQl + swz(tid, c8)) = lo4;
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t t_lane = tmem_base + ((uint32_t)(warp * 32) << 16); // this warp's TMEM lane quarter
  const int lo_min = s_lo_min, hi_max = s_hi_max;

  float o[AT_HD];
#pragma unroll
  for (int i = 0; i < AT_HD; i++)
    o[i] = 0.0f;
  float m_run = -INFINITY, l_run = 0.0f;
  uint32_t ph_s = 0, ph_o = 0;
  const float *kbase = qkv + D + h * AT_HD, *vbase = qkv + 2 * D + h * AT_HD;

  // This thread's share of a K / V block: key (tid / 2), 32 dims (half = tid % 2): 8 float4 each, held in
  // registers so that the loads of block j+1 are in flight while block j goes through its MMA -> softmax
  // -> MMA chain (two co-resident CTAs alone do not hide the L2 round trip of 128 keys x 512 B).
  const int key = tid >> 1, half = tid & 1;
  float4 kr[8], vr[8];
  auto fetch = [&](int jb) {
    const int j = jb + key;
    const bool live = j >= 0 && j < rows && j <= hi_max;
    const float *kp = kbase + (long)(live ? j : 0) * ldq + half * 32;
    const float *vp = vbase + (long)(live ? j : 0) * ldq + half * 32;
#pragma unroll
    for (int i = 0; i < 8; i++) {
      kr[i] = live ? *reinterpret_cast<const float4 *>(kp + i * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
      vr[i] = live ? *reinterpret_cast<const float4 *>(vp + i * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  if (hi_max >= 0) {
    fetch(lo_min);
    for (int jb = lo_min; jb <= hi_max; jb += AT_K) {
      // ---- K and V block: split bf16, swizzled [key][dim] rows (rows outside [0, rows) are zero) ----
#pragma unroll
      for (int i = 0; i < 4; i++) {
        uint4 hi, lo4;
        split8(kr[2 * i], kr[2 * i + 1], hi, lo4);
        *reinterpret_cast<uint4 *>(sKh + swz(key, half * 4 + i)) = hi;
        *reinterpret_cast<uint4 *>(sKl + swz(key, half * 4 + i)) = lo4;
        split8(vr[2 * i], vr[2 * i + 1], hi, lo4);
        *reinterpret_cast<uint4 *>(sVh + swz(key, half * 4 + i)) = hi;
        *reinterpret_cast<uint4 *>(sVl + swz(key, half * 4 + i)) = lo4;
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); // generic writes -> tensor-core reads
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncthreads();
      // ---- S = Q K^T ----
      if (tid == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        mma_split<false>(tmem_base, smem_u32(sQh), smem_u32(sQl), smem_u32(sKh), smem_u32(sKl));
        umma_commit(bar_s);
      }
      if (jb + AT_K <= hi_max)
        fetch(jb + AT_K); // next block's rows: in flight until the top of the next iteration
      mbar_wait(bar_s, ph_s);
      ph_s ^= 1;
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      // ---- mask + online softmax of this thread's row (two passes over its TMEM row: block maximum, then
      //      exponentials 32 columns at a time so only 32 score registers are live); P -> shared memory ----
      float alpha = 1.0f;
      {
        uint32_t sv[32];
        float bm = -INFINITY;
#pragma unroll
        for (int hf = 0; hf < 2; hf++) {
          tmem_ld32(t_lane + (uint32_t)(hf * 32), sv);
#pragma unroll
          for (int c = 0; c < 32; c++) {
            const int j = jb + hf * 32 + c;
            const bool ok = n > 0 && j >= lo && j <= r;
            bm = fmaxf(bm, ok ? __uint_as_float(sv[c]) : -INFINITY);
          }
        }
        const float m_new = fmaxf(m_run, bm);
        // a row that has seen no allowed key yet keeps m = -inf: all its p are 0
        alpha = m_new == -INFINITY ? 1.0f : __expf(m_run - m_new);
        const float sh = m_new == -INFINITY ? 0.0f : m_new;
        m_run = m_new;
        float bs = 0.0f;
#pragma unroll
        for (int hf = 0; hf < 2; hf++) {
          tmem_ld32(t_lane + (uint32_t)(hf * 32), sv);
          float p[32];
#pragma unroll
          for (int c = 0; c < 32; c++) {
            const int j = jb + hf * 32 + c;
            const bool ok = n > 0 && j >= lo && j <= r;
            p[c] = ok ? __expf(__uint_as_float(sv[c]) - sh) : 0.0f;
            bs += p[c];
          }
#pragma unroll
          for (int c8 = 0; c8 < 4; c8++) {
            uint4 hi, lo4;
            split8(make_float4(p[c8 * 8], p[c8 * 8 + 1], p[c8 * 8 + 2], p[c8 * 8 + 3]),
                   make_float4(p[c8 * 8 + 4], p[c8 * 8 + 5], p[c8 * 8 + 6], p[c8 * 8 + 7]), hi, lo4);
            *reinterpret_cast<uint4 *>(sPh + swz(tid, hf * 4 + c8)) = hi;
            *reinterpret_cast<uint4 *>(sPl + swz(tid, hf * 4 + c8)) = lo4;
          }
        }
        l_run = l_run * alpha + bs;
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncthreads();
      // ---- PV = P V ----
      if (tid == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        mma_split<true>(tmem_base + 64, smem_u32(sPh), smem_u32(sPl), smem_u32(sVh), smem_u32(sVl));
        umma_commit(bar_o);
      }
      mbar_wait(bar_o, ph_o);
      ph_o ^= 1;
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
      for (int hf = 0; hf < 2; hf++) {
        uint32_t pv[32];
        tmem_ld32(t_lane + (uint32_t)(64 + hf * 32), pv);
#pragma unroll
        for (int c = 0; c < 32; c++)
          o[hf * 32 + c] = o[hf * 32 + c] * alpha + __uint_as_float(pv[c]);
      }
      // both accumulators and the K / V / P tiles are free again once every thread is past this point
      // (the next iteration's first __syncthreads orders the overwrite after these reads)
    }
  }
  // ---- normalise and store (an all-masked row is zeros) ----
  if (r < rows) {
    const float inv = l_run > 0.0f ? 1.0f / l_run : 0.0f;
    const int col0 = h * AT_HD;
#pragma unroll
    for (int c4 = 0; c4 < AT_HD; c4 += 4) {
      const float4 v = make_float4(o[c4] * inv, o[c4 + 1] * inv, o[c4 + 2] * inv, o[c4 + 3] * inv);
      if (out)
        *reinterpret_cast<float4 *>(out + (long)r * ldo + col0 + c4) = v;
      if (sp.hi)
        split_store4(sp, r, col0 + c4, v);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(128) : "memory");
}

} // namespace

// head_dim 64, ctx > 0.  Returns 0 (nothing launched) when the shape is not supported so the caller
// can take the CUDA-core kernel.
int attention_window_tc(const float *qkv, long ldq, float *out, long ldo, int rows, int D, int n_heads,
                        const int32_t *row_pos, int ctx, cudaStream_t st, SplitOut sp) {
  if (rows <= 0 || ctx <= 0 || D != n_heads * AT_HD || (ldq & 3) != 0)
    return 0;
  static volatile bool configured[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (!configured[dev & 63]) {
    if (cudaFuncSetAttribute(attention_window_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AT_SMEM) != cudaSuccess)
      return 0;
    configured[dev & 63] = true;
  }
  dim3 grid((rows + AT_Q - 1) / AT_Q, n_heads);
  attention_window_tc_kernel<<<grid, AT_THREADS, AT_SMEM, st>>>(qkv, ldq, out, ldo, rows, D, row_pos, ctx, sp);
  return 1;
}

} // namespace ptts
