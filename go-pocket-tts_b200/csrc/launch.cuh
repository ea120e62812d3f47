// launch.cuh — programmatic dependent launch (PDL) for the AR-step kernel chain.
//
// One AR frame is ~79 dependent launches of short kernels; with plain stream order every
// kernel pays the launch latency and its own prologue after the previous grid has fully
// drained.  Kernels on that path start with PDL_TRIGGER (lets the next grid be scheduled
// as SMs free up) and PDL_WAIT (blocks until the previous grid has completed and its
// writes are visible) before they touch any global memory another kernel produces or
// reads; the tcgen05 GEMM runs its prologue and prefetches its (constant) weight tiles
// in front of PDL_WAIT.  Both instructions are no-ops in a launch without the attribute.
#pragma once
#include <cuda_runtime.h>
#include <utility>

namespace ptts {

#define PDL_TRIGGER() asm volatile("griddepcontrol.launch_dependents;" ::: "memory")
#define PDL_WAIT() asm volatile("griddepcontrol.wait;" ::: "memory")

// Set by the engine around the AR step (graph capture keeps the programmatic edges).  Per thread:
// every engine of a multi-GPU pool launches from its own scheduler thread.
bool pdl_enabled();
void pdl_set(bool on);

// cudaFuncSetAttribute is per device: launch wrappers that configure a kernel lazily keep one
// flag per device ordinal (several engines of one process may sit on different GPUs).
// Returns the flag of the current device; the caller sets it AFTER configuring (two engines on one
// device may both configure once, which is harmless; neither launches unconfigured).
inline volatile bool &device_flag(volatile bool (&flags)[64]) {
  int dev = 0;
  cudaGetDevice(&dev);
  return flags[dev & 63];
}

// cluster > 1: thread-block clusters of that many consecutive CTAs along x (grid.x must be a
// multiple of it).
template <typename... KArgs, typename... Args>
inline cudaError_t launch_kc(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, int cluster,
                             Args &&...args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (pdl_enabled()) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    na++;
  }
  if (cluster > 1) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = (unsigned)cluster;
    attr[na].val.clusterDim.y = 1;
    attr[na].val.clusterDim.z = 1;
    na++;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}
template <typename... KArgs, typename... Args>
inline cudaError_t launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                            Args &&...args) {
  return launch_kc(kernel, grid, block, smem, st, 1, std::forward<Args>(args)...);
}

} // namespace ptts
