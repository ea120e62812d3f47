// flow_head.cuh — parameter block of the persistent batched flow-head kernel (flow_head.cu).
#pragma once
#include "kernels.cuh"
#include <cuda.h>

namespace ptts {

constexpr int FH_MAX_PHASES = 48;
constexpr size_t FH_BARRIER_BYTES = 16 * 32 * 4; // 16 striped counters, 128 B apart

enum FhKind : int { FH_GEMM = 0, FH_ROWS_HEAD = 1, FH_ROWS_LN = 2 };

// One tensor-core GEMM of the chain: C = epi(A [B, K] x W [N, K]^T), both operands behind TMA maps
// that live in device memory (128-byte CUtensorMap blobs, 64-byte aligned).
struct FhGemm {
  const CUtensorMap *wmap, *amap;
  int n_rows_hi; // rows of the weight's hi plane (= N of the weight matrix)
  int a_lo_row;  // row offset of A's lo plane inside amap
  int N, K, bn;
  int epi_kind;  // epi_code() of a compiled epilogue, -1: generic
  GemmEpi epi;
};

struct FhPhase {
  int kind;
  int no_barrier_after;
  // FH_ROWS_LN: h = (LN(xf) * ln_w + ln_b) * (1 + ada[:, scale_off ..]) + ada[:, shift_off ..]
  const float *ln_w, *ln_b;
  float ln_eps;
  int scale_off, shift_off;
  FhGemm g;
};

struct FhParams {
  int D, C, ldim, ada_n, n_phases;
  int debug; // timing experiments only (PTTS_FH_DEBUG): 1 skip grid barriers, 2 skip GEMM items, 4 skip row phases
  // inputs / outputs (fp32)
  const float *x;     // [B, D] transformer output
  const float *noise; // [B, ldim]
  float *last;        // [B, D] out_norm(x) (parity tap)
  float *eos;         // [B]
  float *cur;         // [B, ldim] latent being integrated (output)
  float *xf;          // [B, C] flow-net residual stream
  float *ada;         // [B, ada_n]
  // split planes of the GEMM inputs
  SplitOut s_last, s_cur, s_h;
  // row-wise weights
  const float *on_w, *on_b;
  float on_eps;
  const float *eos_w;
  float eos_b;
  const float *ip_w, *ip_b;
  unsigned *barrier;
  int *error;
  FhPhase phase[FH_MAX_PHASES];
};

size_t flow_head_smem();
int flow_head_grid(); // CTAs of the cooperative launch (0: the kernel cannot be resident)
int flow_head_epi_kind(const GemmEpi &e);
// Launches barrier reset + kernel; returns the number of launches (2) or -1.
int flow_head_launch(const FhParams *dp, int grid, int B, unsigned *barrier, cudaStream_t st);

} // namespace ptts
