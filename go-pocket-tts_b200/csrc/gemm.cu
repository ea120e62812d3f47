// gemm.cu — fp32 "TN" GEMMs (C = A * W^T, both K-contiguous) with the fused
// epilogue of kernels.cuh.  Replaces the reference's per-output DotProduct loop
// (internal/native/linear.go:117-182 over tensor/dot_amd64.s:34-116) and, through
// overlapping A rows, its im2col Conv1D / dot-scatter ConvTranspose1D
// (ops/conv1d.go:20-83, ops/convtranspose1d.go:73-148).
//
// Two regimes:
//  * M <= 8   : weight-streaming skinny kernel.  HBM-bound: every weight byte is
//               read exactly once with 128-bit non-allocating loads, activations
//               come from L1/L2, warp-shuffle reduction, epilogue on lane 0.
//  * M  > 8   : register-tiled SGEMM (CUDA cores, fp32 FMA, fp32 accumulate),
//               double-buffered shared-memory tiles.  (The tcgen05 split-bf16
//               path replaces this for the batched regime; see DESIGN.md.)
#include "kernels.cuh"
#include "gemm_epi.cuh"
#include "launch.cuh"
#include <math.h>

namespace ptts {

__device__ __forceinline__ float4 ldg_stream(const float *p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}

// ---------------------------------------------------------------------------
// skinny kernel: MR rows (<= 8), NC output columns per warp pass
// ---------------------------------------------------------------------------
// KU = float4 weight loads per lane and column issued before the first FMA: the
// kernel is HBM-bound and a lane that waits on one or two loads at a time leaves the
// memory system idle (B=1 ran at 0.7 TB/s that way).  K <= 128*KU is a single trip.
template <int MR, int NC, int KU>
__global__ void __launch_bounds__(256) gemv_kernel(GemmArgs g) {
  PDL_TRIGGER();
  PDL_WAIT();
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  const int M = g.M, N = g.N, K = g.K;
  for (int n0 = warp * NC; n0 < N; n0 += nwarps * NC) {
    float acc[MR][NC];
#pragma unroll
    for (int r = 0; r < MR; r++)
#pragma unroll
      for (int c = 0; c < NC; c++)
        acc[r][c] = 0.0f;
    for (int k0 = lane * 4; k0 < K; k0 += 128 * KU) {
      float4 w[KU][NC];
#pragma unroll
      for (int u = 0; u < KU; u++)
#pragma unroll
        for (int c = 0; c < NC; c++) {
          const int n = n0 + c < N ? n0 + c : N - 1;
          const int k = k0 + u * 128;
          w[u][c] = k < K ? ldg_stream(g.W + (long)n * g.ldw + k) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
      for (int u = 0; u < KU; u++) {
        const int k = k0 + u * 128;
        if (k < K) {
#pragma unroll
          for (int r = 0; r < MR; r++) {
            const int rr = r < M ? r : M - 1;
            const float4 x = *reinterpret_cast<const float4 *>(g.A + (long)rr * g.lda + k);
#pragma unroll
            for (int c = 0; c < NC; c++) {
              acc[r][c] = fmaf(x.x, w[u][c].x, acc[r][c]);
              acc[r][c] = fmaf(x.y, w[u][c].y, acc[r][c]);
              acc[r][c] = fmaf(x.z, w[u][c].z, acc[r][c]);
              acc[r][c] = fmaf(x.w, w[u][c].w, acc[r][c]);
            }
          }
        }
      }
    }
#pragma unroll
    for (int r = 0; r < MR; r++)
#pragma unroll
      for (int c = 0; c < NC; c++) {
        float v = acc[r][c];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
          v += __shfl_xor_sync(0xffffffffu, v, o);
        acc[r][c] = v;
      }
    // lanes 0..MR*NC-1 each finish one output (spreads the epilogue loads)
    if (lane < MR * NC) {
      int r = lane / NC, c = lane % NC;
      float v = 0.0f;
#pragma unroll
      for (int rr = 0; rr < MR; rr++)
#pragma unroll
        for (int cc = 0; cc < NC; cc++)
          if (rr == r && cc == c)
            v = acc[rr][cc];
      int n = n0 + c;
      if (r < M && n < N) {
        const GemmEpi &e = g.epi;
        v = epi_one(e, v, r, n);
        if (e.rowmask && !e.rowmask[r / e.mask_div])
          v = 0.0f;
        if (e.C)
          e.C[(long)r * e.ldc + n] = v;
        if (e.C2)
          e.C2[(long)r * e.ldc2 + n] = act_apply(v, ACT_ELU);
      }
    }
  }
}

// ---------------------------------------------------------------------------
// tiled SGEMM
// ---------------------------------------------------------------------------
// Split-K (gridDim.z > 1): every CTA writes its partial tile to the workspace,
// the last CTA to arrive on the tile's counter sums the partials in split order
// (deterministic) and runs the epilogue; the counter resets itself.
template <int BM, int BN, int RM, int RN>
__global__ void __launch_bounds__((BM / (4 * RM)) * (BN / (4 * RN)),
                                  ((BM / (4 * RM)) * (BN / (4 * RN)) >= 256 ? 2 : 3))
    sgemm_tn_kernel(GemmArgs g) {
  constexpr int BK = 16;
  constexpr int TX = BN / (4 * RN), TY = BM / (4 * RM), NT = TX * TY;
  constexpr int LDA_S = BM + 4, LDB_S = BN + 4;
  constexpr int A_F4 = BM * BK / 4, B_F4 = BN * BK / 4;
  constexpr int A_PER = (A_F4 + NT - 1) / NT, B_PER = (B_F4 + NT - 1) / NT;
  __shared__ __align__(16) float As[2][BK][LDA_S];
  __shared__ __align__(16) float Bs[2][BK][LDB_S];

  const int tid = threadIdx.x;
  const int tx = tid % TX, ty = tid / TX;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int M = g.M, N = g.N, K = g.K;

  float4 ra[A_PER], rb[B_PER];
  auto gload = [&](int k0) {
#pragma unroll
    for (int i = 0; i < A_PER; i++) {
      int idx = tid + i * NT;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (idx < A_F4) {
        int row = idx >> 2, kq = idx & 3;
        int gm = m0 + row, gk = k0 + kq * 4;
        if (gm < M && gk < K)
          v = *reinterpret_cast<const float4 *>(g.A + (long)gm * g.lda + gk);
      }
      ra[i] = v;
    }
#pragma unroll
    for (int i = 0; i < B_PER; i++) {
      int idx = tid + i * NT;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (idx < B_F4) {
        int row = idx >> 2, kq = idx & 3;
        int gn = n0 + row, gk = k0 + kq * 4;
        if (gn < N && gk < K)
          v = __ldg(reinterpret_cast<const float4 *>(g.W + (long)gn * g.ldw + gk));
      }
      rb[i] = v;
    }
  };
  auto sstore = [&](int buf) {
#pragma unroll
    for (int i = 0; i < A_PER; i++) {
      int idx = tid + i * NT;
      if (idx < A_F4) {
        int row = idx >> 2, kq = (idx & 3) * 4;
        As[buf][kq + 0][row] = ra[i].x;
        As[buf][kq + 1][row] = ra[i].y;
        As[buf][kq + 2][row] = ra[i].z;
        As[buf][kq + 3][row] = ra[i].w;
      }
    }
#pragma unroll
    for (int i = 0; i < B_PER; i++) {
      int idx = tid + i * NT;
      if (idx < B_F4) {
        int row = idx >> 2, kq = (idx & 3) * 4;
        Bs[buf][kq + 0][row] = rb[i].x;
        Bs[buf][kq + 1][row] = rb[i].y;
        Bs[buf][kq + 2][row] = rb[i].z;
        Bs[buf][kq + 3][row] = rb[i].w;
      }
    }
  };

  float acc[4 * RM][4 * RN];
#pragma unroll
  for (int i = 0; i < 4 * RM; i++)
#pragma unroll
    for (int j = 0; j < 4 * RN; j++)
      acc[i][j] = 0.0f;

  const int nk_all = (K + BK - 1) / BK;
  const int S = gridDim.z;
  const int kt_per = (nk_all + S - 1) / S;
  const int kt0 = blockIdx.z * kt_per;
  const int nk = min(nk_all - kt0, kt_per) > 0 ? min(nk_all - kt0, kt_per) : 0;
  if (nk > 0) {
    gload(kt0 * BK);
    sstore(0);
  }
  __syncthreads();
  for (int kt = 0; kt < nk; kt++) {
    const int buf = kt & 1;
    if (kt + 1 < nk)
      gload((kt0 + kt + 1) * BK);
#pragma unroll
    for (int k = 0; k < BK; k++) {
      float a[4 * RM], b[4 * RN];
#pragma unroll
      for (int rm = 0; rm < RM; rm++) {
        float4 t = *reinterpret_cast<const float4 *>(&As[buf][k][rm * (BM / RM) + ty * 4]);
        a[rm * 4 + 0] = t.x; a[rm * 4 + 1] = t.y; a[rm * 4 + 2] = t.z; a[rm * 4 + 3] = t.w;
      }
#pragma unroll
      for (int rn = 0; rn < RN; rn++) {
        float4 t = *reinterpret_cast<const float4 *>(&Bs[buf][k][rn * (BN / RN) + tx * 4]);
        b[rn * 4 + 0] = t.x; b[rn * 4 + 1] = t.y; b[rn * 4 + 2] = t.z; b[rn * 4 + 3] = t.w;
      }
#pragma unroll
      for (int i = 0; i < 4 * RM; i++)
#pragma unroll
        for (int j = 0; j < 4 * RN; j++)
          acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (kt + 1 < nk)
      sstore(buf ^ 1);
    __syncthreads();
  }

  if (S > 1) {
    __shared__ int s_last;
    const int tile = blockIdx.y * gridDim.x + blockIdx.x;
    constexpr int FR = 4 * RM * RN; // float4 fragments per thread
    float4 *wsp = reinterpret_cast<float4 *>(g.ws) + ((long)tile * S) * FR * NT;
    float4 *mine = wsp + (long)blockIdx.z * FR * NT;
#pragma unroll
    for (int i = 0; i < 4 * RM; i++)
#pragma unroll
      for (int rn = 0; rn < RN; rn++)
        mine[(i * RN + rn) * NT + tid] = make_float4(acc[i][rn * 4 + 0], acc[i][rn * 4 + 1],
                                                     acc[i][rn * 4 + 2], acc[i][rn * 4 + 3]);
    __threadfence();
    __syncthreads();
    if (tid == 0) {
      int prev = atomicAdd(g.counters + tile, 1);
      s_last = (prev == S - 1);
      if (s_last)
        g.counters[tile] = 0;
    }
    __syncthreads();
    if (!s_last)
      return;
    __threadfence();
#pragma unroll
    for (int i = 0; i < 4 * RM; i++)
#pragma unroll
      for (int j = 0; j < 4 * RN; j++)
        acc[i][j] = 0.0f;
    // all FR fragments of one split are loaded back-to-back (FR independent
    // L2 loads in flight) before they are accumulated
    for (int sp = 0; sp < S; sp++) {
      float4 u[FR];
#pragma unroll
      for (int f = 0; f < FR; f++)
        u[f] = __ldcg(wsp + ((long)sp * FR + f) * NT + tid);
#pragma unroll
      for (int i = 0; i < 4 * RM; i++)
#pragma unroll
        for (int rn = 0; rn < RN; rn++) {
          const float4 t = u[i * RN + rn];
          acc[i][rn * 4 + 0] += t.x; acc[i][rn * 4 + 1] += t.y;
          acc[i][rn * 4 + 2] += t.z; acc[i][rn * 4 + 3] += t.w;
        }
    }
  }

  const GemmEpi &e = g.epi;
#pragma unroll
  for (int rm = 0; rm < RM; rm++)
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const int m = m0 + rm * (BM / RM) + ty * 4 + i;
      if (m >= M)
        continue;
      const bool live = !e.rowmask || e.rowmask[m / e.mask_div];
#pragma unroll
      for (int rn = 0; rn < RN; rn++) {
        const int n = n0 + rn * (BN / RN) + tx * 4;
        if (n >= N)
          continue;
        float4 v = epi_vec4(e, make_float4(acc[rm * 4 + i][rn * 4 + 0], acc[rm * 4 + i][rn * 4 + 1],
                                           acc[rm * 4 + i][rn * 4 + 2], acc[rm * 4 + i][rn * 4 + 3]), m, n);
        if (!live)
          v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (e.C)
          *reinterpret_cast<float4 *>(e.C + (long)m * e.ldc + n) = v;
        if (e.C2) {
          float4 u;
          u.x = act_apply(v.x, ACT_ELU);
          u.y = act_apply(v.y, ACT_ELU);
          u.z = act_apply(v.z, ACT_ELU);
          u.w = act_apply(v.w, ACT_ELU);
          *reinterpret_cast<float4 *>(e.C2 + (long)m * e.ldc2 + n) = u;
        }
      }
    }
}

template <int BM, int BN, int RM, int RN>
static void launch_sgemm(const GemmArgs &g, cudaStream_t st) {
  dim3 grid((g.N + BN - 1) / BN, (g.M + BM - 1) / BM, 1);
  constexpr int NT = (BM / (4 * RM)) * (BN / (4 * RN));
  // split-K until ~2 CTAs per SM, bounded by the K loop and the workspace
  const long tiles = (long)grid.x * grid.y;
  const int nk = (g.K + 15) / 16;
  int S = 1;
  if (g.ws && g.counters && tiles < 148 && tiles <= g.n_counters) {
    S = (int)((2 * 148 + tiles - 1) / tiles);
    if (S > nk / 4) S = nk / 4;
    if (S > 32) S = 32;
    while (S > 1 && tiles * S * (long)BM * BN > g.ws_floats) S--;
    if (S < 1) S = 1;
  }
  grid.z = S;
  sgemm_tn_kernel<BM, BN, RM, RN><<<grid, NT, 0, st>>>(g);
}

template <int MR>
static void launch_gemv(const GemmArgs &g, cudaStream_t st) {
  // one column per warp pass keeps every SM streaming for N ~ 1e3; every warp issues
  // all the weight loads of its pass (up to 2 KB per lane) before it consumes any
  const int warps_per_cta = 8;
  if (g.N >= 8192) {
    int ctas = (g.N / 2 + warps_per_cta - 1) / warps_per_cta;
    if (ctas > 148 * 8)
      ctas = 148 * 8;
    launch_k(gemv_kernel<MR, 2, 4>, dim3(ctas), dim3(warps_per_cta * 32), 0, st, g);
  } else {
    int ctas = (g.N + warps_per_cta - 1) / warps_per_cta;
    if (ctas > 148 * 8)
      ctas = 148 * 8;
    if (g.K <= 512)
      launch_k(gemv_kernel<MR, 1, 4>, dim3(ctas), dim3(warps_per_cta * 32), 0, st, g);
    else if (g.K <= 1024 || MR > 2)
      launch_k(gemv_kernel<MR, 1, 8>, dim3(ctas), dim3(warps_per_cta * 32), 0, st, g);
    else
      launch_k(gemv_kernel<MR, 1, 16>, dim3(ctas), dim3(warps_per_cta * 32), 0, st, g);
  }
}

static thread_local const char *g_variant = "";
const char *gemm_last_variant() { return g_variant; }

int gemm(const GemmArgs &g, cudaStream_t st) {
  if (g.M <= 0 || g.N <= 0)
    return 0;
  if (g.M <= 8) {
    g_variant = "gemv_f32";
    if (g.M == 1)
      launch_gemv<1>(g, st);
    else if (g.M == 2)
      launch_gemv<2>(g, st);
    else if (g.M <= 4)
      launch_gemv<4>(g, st);
    else
      launch_gemv<8>(g, st);
    return 1;
  }
  auto ctas = [&](int bm, int bn) {
    return (long)((g.M + bm - 1) / bm) * ((g.N + bn - 1) / bn);
  };
  const long want = 148; // at least one CTA per SM before growing the tile
  const bool can_split = g.ws && g.counters && g.K >= 256;
  if (can_split && g.N >= 128 && g.M <= 512) {
    // skinny-M regime of the batched AR step: big tiles + split-K keep all SMs
    // streaming weights
    if (g.M > 64) {
      g_variant = "sgemm_128x128";
      launch_sgemm<128, 128, 2, 2>(g, st);
    } else {
      g_variant = "sgemm_64x128";
      launch_sgemm<64, 128, 2, 2>(g, st);
    }
    return 1;
  }
  if (g.N <= 32) {
    if (ctas(128, 32) >= want) {
      g_variant = "sgemm_128x32";
      launch_sgemm<128, 32, 2, 1>(g, st);
    } else {
      g_variant = "sgemm_64x32";
      launch_sgemm<64, 32, 2, 1>(g, st);
    }
  } else if (ctas(128, 128) >= want && g.N >= 128) {
    g_variant = "sgemm_128x128";
    launch_sgemm<128, 128, 2, 2>(g, st);
  } else if (ctas(64, 128) >= want && g.N >= 128) {
    g_variant = "sgemm_64x128";
    launch_sgemm<64, 128, 2, 2>(g, st);
  } else if (ctas(64, 64) >= want || g.M > 32) {
    g_variant = "sgemm_64x64";
    launch_sgemm<64, 64, 2, 2>(g, st);
  } else {
    g_variant = "sgemm_32x64";
    launch_sgemm<32, 64, 1, 2>(g, st);
  }
  return 1;
}

__global__ void __launch_bounds__(256)
    rowdot_kernel(const float *A, long lda, const float *w, float bias0, int M,
                  int K, float *out) {
  PDL_TRIGGER();
  PDL_WAIT();
  const int lane = threadIdx.x & 31;
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (row >= M)
    return;
  const float *a = A + (long)row * lda;
  float s = 0.0f;
  for (int k = lane * 4; k < K; k += 128) {
    float4 x = *reinterpret_cast<const float4 *>(a + k);
    float4 y = __ldg(reinterpret_cast<const float4 *>(w + k));
    s = fmaf(x.x, y.x, s);
    s = fmaf(x.y, y.y, s);
    s = fmaf(x.z, y.z, s);
    s = fmaf(x.w, y.w, s);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
    s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0)
    out[row] = s + bias0;
}

int rowdot(const float *A, long lda, const float *w, float bias0, int M, int K,
           float *out, cudaStream_t st) {
  if (M <= 0)
    return 0;
  int ctas = (M + 7) / 8;
  launch_k(rowdot_kernel, dim3(ctas), dim3(256), 0, st, A, lda, w, bias0, M, K, out);
  return 1;
}

} // namespace ptts
