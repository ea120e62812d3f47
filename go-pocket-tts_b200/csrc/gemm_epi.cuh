// gemm_epi.cuh — the fused GEMM epilogue shared by the CUDA-core and the
// tcgen05 kernels (order documented at GemmEpi in kernels.cuh).
#pragma once
#include "kernels.cuh"
#include <math.h>

namespace ptts {

__device__ __forceinline__ float act_apply(float v, int act) {
  switch (act) {
  case ACT_GELU: // exact-erf GELU (internal/native/tensor_util.go:84-94)
    return 0.5f * v * (1.0f + erff(v * 0.70710678118654752440f));
  case ACT_SILU: // tensor_util.go:73-82
    return v / (1.0f + expf(-v));
  case ACT_ELU: // tensor_util.go:119-128
    return v > 0.0f ? v : expf(v) - 1.0f;
  default:
    return v;
  }
}

__device__ __forceinline__ float epi_one(const GemmEpi &e, float acc, int m, int n) {
  float v = acc;
  if (e.bias)
    v += __ldg(e.bias + n);
  v = act_apply(v, e.act);
  if (e.colscale)
    v *= __ldg(e.colscale + n);
  if (e.mul)
    v *= e.mul[(long)m * e.ldmul + n];
  v *= e.alpha;
  if (e.residual)
    v += e.residual[(long)m * e.ldr + n];
  return v;
}

// Four consecutive columns at once (n % 4 == 0, all operand rows 16-byte aligned):
// vector loads and warp-uniform branches instead of four scalar epi_one calls.
__device__ __forceinline__ float4 epi_vec4(const GemmEpi &e, float4 a, int m, int n) {
  if (e.bias) {
    const float4 b = __ldg(reinterpret_cast<const float4 *>(e.bias + n));
    a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
  }
  if (e.act != ACT_NONE) {
    a.x = act_apply(a.x, e.act); a.y = act_apply(a.y, e.act);
    a.z = act_apply(a.z, e.act); a.w = act_apply(a.w, e.act);
  }
  if (e.colscale) {
    const float4 c = __ldg(reinterpret_cast<const float4 *>(e.colscale + n));
    a.x *= c.x; a.y *= c.y; a.z *= c.z; a.w *= c.w;
  }
  if (e.mul) {
    const float4 t = *reinterpret_cast<const float4 *>(e.mul + (long)m * e.ldmul + n);
    a.x *= t.x; a.y *= t.y; a.z *= t.z; a.w *= t.w;
  }
  if (e.alpha != 1.0f) {
    a.x *= e.alpha; a.y *= e.alpha; a.z *= e.alpha; a.w *= e.alpha;
  }
  if (e.residual) {
    const float4 r = *reinterpret_cast<const float4 *>(e.residual + (long)m * e.ldr + n);
    a.x += r.x; a.y += r.y; a.z += r.z; a.w += r.w;
  }
  return a;
}

} // namespace ptts
