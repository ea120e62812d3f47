// tc_common.cuh — tcgen05 / TMEM / TMA / mbarrier primitives shared by the GEMM kernel
// (tcgemm.cu) and the fused SEANet residual-block kernel (resblock.cu).
#pragma once
#include "kernels.cuh"
#include <cuda.h>
#include <cuda_bf16.h>
#include "gemm_epi.cuh"
#include "split.cuh"

namespace ptts {
namespace {

constexpr int TC_BM = 128;
constexpr int TC_BK = 64; // bf16 elements per k-block = one 128-byte swizzle row
constexpr int TC_CONV_THREADS = 256;
// threads: warp 0 TMA, warp 1 MMA, EW epilogue warps (4, or 8 = two per TMEM lane quarter
// for the short-K shapes whose epilogue is the bottleneck), then 8 A-converter warps when
// A is fp32 in HBM.
__host__ __device__ constexpr int tc_base_threads(int ew) { return 64 + ew * 32; }
__host__ __device__ constexpr int tc_threads(int ew, bool atma) { return tc_base_threads(ew) + (atma ? 0 : TC_CONV_THREADS); }

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t"
      "}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
               : "memory");
}

__device__ __forceinline__ void tma_load_3d(void *dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

// K-major, 128B-swizzled operand tile: rows of 128 bytes, 8-row atoms 1024 B apart.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);        // start address
  d |= (uint64_t)0 << 16;                          // leading byte offset (unused, swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;                // stride byte offset: 8 rows * 128 B
  d |= (uint64_t)1 << 46;                          // descriptor version (sm_100)
  d |= (uint64_t)2 << 61;                          // SWIZZLE_128B
  return d;
}
// kind::f16 instruction descriptor: D=f32, A=B=bf16, both K-major, M=128, N=BN.
__host__ __device__ constexpr uint32_t make_idesc(int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ float4 lds128(uint32_t addr) {
  float4 r;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "r"(addr) : "memory");
  return r;
}


// Fused epilogue of one 32x32 accumulator chunk held in a warp's staging buffer: lane
// (rsub, cf) owns columns n..n+3 of rows rsub, rsub+4, ... (8 rows).  `rows` bit j = row
// in range, bit 8+j = row live (not masked).  Outputs (each optional): C fp32, Cs its
// bf16 split, C2 = elu(C) fp32, C2s its split.
//
// The hot configurations are compiled as straight-line code (activation ACT, output set
// MASK: 1 = C, 2 = Cs, 4 = C2, 8 = C2s, FULL = all 8 rows in range); anything else (adaLN
// gate `mul`, `alpha`, unusual output sets) takes the generic path with runtime checks.
// ELU with the hardware exponential: abs error ~1e-7 on values in (-1, 0], far inside the waveform tolerance;
// the full-precision expf was most of the epilogue's instructions on the HBM-bound SEANet layers.
// exp is ex2.approx.ftz(v * log2 e): __expf is the same product and the same MUFU, but its non-ftz
// ex2 wraps the MUFU in denormal scaling (two more FMUL and an FSETP per element — a fifth of the instructions of
// the conv epilogues, which are what bounds the SEANet kernels).  For v <= 0 the results differ only where
// exp(v) or v itself is denormal: exp(v) - 1 is -1 either way, and an input below 1e-38 gives 0 instead of v.
__device__ __forceinline__ float elu_fast(float v) {
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(__fmul_rn(v, 1.4426950216293335f)));
  return v > 0.0f ? v : e - 1.0f;
}
__device__ __forceinline__ float4 elu4(const float4 &o) {
  return make_float4(elu_fast(o.x), elu_fast(o.y), elu_fast(o.z), elu_fast(o.w));
}
template <int ACT, int MASK, bool FULL>
__device__ __forceinline__ void epi_rows_fast(const GemmEpi &e, uint32_t stg_addr, int stg_ld_bytes, long m_first,
                                              int n, uint32_t rows, const float4 &bias4, const float4 &cs4,
                                              const float4 *res, bool have_res) {
  float *cp = (MASK & 1) ? e.C + m_first * e.ldc + n : nullptr;
  float *cp2 = (MASK & 4) ? e.C2 + m_first * e.ldc2 + n : nullptr;
#pragma unroll
  for (int j = 0; j < 8; j++) {
    if (FULL || ((rows >> j) & 1)) {
      float4 o = lds128(stg_addr + j * 4 * stg_ld_bytes);
      o.x += bias4.x; o.y += bias4.y; o.z += bias4.z; o.w += bias4.w;
      if (ACT != ACT_NONE) {
        o.x = act_apply(o.x, ACT); o.y = act_apply(o.y, ACT);
        o.z = act_apply(o.z, ACT); o.w = act_apply(o.w, ACT);
      }
      o.x *= cs4.x; o.y *= cs4.y; o.z *= cs4.z; o.w *= cs4.w;
      if (have_res) {
        o.x += res[j].x; o.y += res[j].y; o.z += res[j].z; o.w += res[j].w;
      }
      if (!((rows >> (8 + j)) & 1))
        o = make_float4(0.f, 0.f, 0.f, 0.f);
      if (MASK & 1)
        *reinterpret_cast<float4 *>(cp + (long)j * 4 * e.ldc) = o;
      if (MASK & 2)
        split_store4(e.Cs, m_first + j * 4, n, o);
      if (MASK & 12) {
        const float4 u = elu4(o);
        if (MASK & 4)
          *reinterpret_cast<float4 *>(cp2 + (long)j * 4 * e.ldc2) = u;
        if (MASK & 8)
          split_store4(e.C2s, m_first + j * 4, n, u);
      }
    }
  }
}
__device__ __forceinline__ void epi_rows_generic(const GemmEpi &e, uint32_t stg_addr, int stg_ld_bytes, long m_first, int n,
                                              uint32_t rows, const float4 &bias4, const float4 &cs4, const float4 *res,
                                              bool have_res) {
#pragma unroll
  for (int j = 0; j < 8; j++) {
    if ((rows >> j) & 1) {
      const long m = m_first + j * 4;
      float4 o = lds128(stg_addr + j * 4 * stg_ld_bytes);
      o.x += bias4.x; o.y += bias4.y; o.z += bias4.z; o.w += bias4.w;
      if (e.act != ACT_NONE) {
        o.x = act_apply(o.x, e.act); o.y = act_apply(o.y, e.act);
        o.z = act_apply(o.z, e.act); o.w = act_apply(o.w, e.act);
      }
      o.x *= cs4.x; o.y *= cs4.y; o.z *= cs4.z; o.w *= cs4.w;
      if (e.mul) {
        const float4 t = *reinterpret_cast<const float4 *>(e.mul + m * e.ldmul + n);
        o.x *= t.x; o.y *= t.y; o.z *= t.z; o.w *= t.w;
      }
      o.x *= e.alpha; o.y *= e.alpha; o.z *= e.alpha; o.w *= e.alpha;
      if (have_res) {
        o.x += res[j].x; o.y += res[j].y; o.z += res[j].z; o.w += res[j].w;
      }
      if (!((rows >> (8 + j)) & 1))
        o = make_float4(0.f, 0.f, 0.f, 0.f);
      if (e.C)
        *reinterpret_cast<float4 *>(e.C + m * e.ldc + n) = o;
      if (e.Cs.hi)
        split_store4(e.Cs, m, n, o);
      if (e.C2 || e.C2s.hi) {
        const float4 u = elu4(o);
        if (e.C2)
          *reinterpret_cast<float4 *>(e.C2 + m * e.ldc2 + n) = u;
        if (e.C2s.hi)
          split_store4(e.C2s, m, n, u);
      }
    }
  }
}
// Epilogue kinds with a compiled fast path: (activation << 4) | output mask.
__host__ __device__ constexpr int epi_code(int act, int mask) { return (act << 4) | mask; }
__host__ inline int epi_kind_of(const GemmEpi &e) {
  if (e.mul || e.alpha != 1.0f)
    return -1;
  const int mask = (e.C ? 1 : 0) | (e.Cs.hi ? 2 : 0) | (e.C2 ? 4 : 0) | (e.C2s.hi ? 8 : 0);
  const int code = epi_code(e.act, mask);
  switch (code) {
  case epi_code(ACT_NONE, 1): case epi_code(ACT_NONE, 3): case epi_code(ACT_NONE, 8): case epi_code(ACT_NONE, 9):
  case epi_code(ACT_NONE, 4): case epi_code(ACT_NONE, 5): case epi_code(ACT_GELU, 1): case epi_code(ACT_GELU, 2):
  case epi_code(ACT_GELU, 3): case epi_code(ACT_SILU, 3):
    return code;
  default:
    return -1;
  }
}
#define TC_EPI_CASE(ACT, MASK)                                                                                         \
  case epi_code(ACT, MASK):                                                                                            \
    if (full)                                                                                                          \
      epi_rows_fast<ACT, MASK, true>(e, stg_addr, stg_ld_bytes, m_first, n, rows, bias4, cs4, res, have_res);          \
    else                                                                                                               \
      epi_rows_fast<ACT, MASK, false>(e, stg_addr, stg_ld_bytes, m_first, n, rows, bias4, cs4, res, have_res);         \
    break;
__device__ __forceinline__ void epi_rows_dispatch(int kind, const GemmEpi &e, uint32_t stg_addr, int stg_ld_bytes,
                                                  long m_first, int n, uint32_t rows, const float4 &bias4,
                                                  const float4 &cs4, const float4 *res, bool have_res, bool full) {
  switch (kind) {
    TC_EPI_CASE(ACT_NONE, 1)
    TC_EPI_CASE(ACT_NONE, 3)
    TC_EPI_CASE(ACT_NONE, 8)
    TC_EPI_CASE(ACT_NONE, 9)
    TC_EPI_CASE(ACT_NONE, 4)
    TC_EPI_CASE(ACT_NONE, 5)
    TC_EPI_CASE(ACT_GELU, 1)
    TC_EPI_CASE(ACT_GELU, 2)
    TC_EPI_CASE(ACT_GELU, 3)
    TC_EPI_CASE(ACT_SILU, 3)
  default:
    epi_rows_generic(e, stg_addr, stg_ld_bytes, m_first, n, rows, bias4, cs4, res, have_res);
    break;
  }
}


} // namespace
} // namespace ptts
