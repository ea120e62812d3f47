// split.cuh — fp32 -> (hi, lo) bf16 split used for the tensor-core operands.
//   x = hi + lo,  hi = bf16_rn(x),  lo = bf16_rn(x - hi)   (~16 mantissa bits)
// Producers of a GEMM's A operand (LayerNorm, attention, activations, GEMM epilogues)
// write the two planes directly so the GEMM can fetch both operands with TMA.
#pragma once
#include "kernels.cuh"

namespace ptts {

// Packed conversions (cvt.rn.bf16x2.f32 -> one F2FP per pair); x0 lands in the low half.
__device__ __forceinline__ void split2(float x0, float x1, uint32_t &hi, uint32_t &lo) {
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(x1), "f"(x0));
  const float h0 = __uint_as_float(hi << 16), h1 = __uint_as_float(hi & 0xFFFF0000u);
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(x1 - h1), "f"(x0 - h0));
}
__device__ __forceinline__ void split8(const float4 &u, const float4 &v, uint4 &hi, uint4 &lo) {
  split2(u.x, u.y, hi.x, lo.x);
  split2(u.z, u.w, hi.y, lo.y);
  split2(v.x, v.y, hi.z, lo.z);
  split2(v.z, v.w, hi.w, lo.w);
}
// Four consecutive columns (col % 4 == 0) of row `row` into both planes.
__device__ __forceinline__ void split_store4(const SplitOut &s, long row, long col, const float4 &v) {
  uint2 hi, lo;
  split2(v.x, v.y, hi.x, lo.x);
  split2(v.z, v.w, hi.y, lo.y);
  uint16_t *p = s.hi + row * s.ld + col;
  *reinterpret_cast<uint2 *>(p) = hi;
  *reinterpret_cast<uint2 *>(p + s.lo_off) = lo;
}
__device__ __forceinline__ void split_store1(const SplitOut &s, long row, long col, float v) {
  uint32_t hi, lo;
  split2(v, 0.0f, hi, lo);
  uint16_t *p = s.hi + row * s.ld + col;
  p[0] = (uint16_t)(hi & 0xFFFFu);
  p[s.lo_off] = (uint16_t)(lo & 0xFFFFu);
}

} // namespace ptts
