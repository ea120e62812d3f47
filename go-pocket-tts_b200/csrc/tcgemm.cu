// tcgemm.cu — tcgen05 / TMEM / TMA GEMM with split-bf16 operands (sm_100a).
//
//   C[M,N] = epi( A[M,K] (fp32, row stride lda)  x  W[N,K]^T )
//
// The reference computes every contraction in fp32 (internal/native/linear.go:
// 117-182) and feeds the result back through up to 234 AR steps, so plain bf16
// operands do not hold the 1e-3 latent tolerance.  Each operand is therefore
// split x = hi + lo (two bf16 values, ~16 mantissa bits) and every k-block issues
// three tensor-core products accumulated in fp32 TMEM:
//     D += Ahi*Whi + Ahi*Wlo + Alo*Whi          (dropped Alo*Wlo ~ 2^-16 relative)
//
//   * W is split ONCE at upload into a bf16 [2N, K] matrix (hi rows then lo
//     rows) and streamed by TMA (cp.async.bulk.tensor, 128B swizzle) — one box
//     for the hi tile, one for the lo tile;
//   * A stays fp32 in HBM (no layout change anywhere else in the engine, and the
//     overlapping-row implicit-GEMM conv trick keeps working): 8 converter warps
//     load it with coalesced 128-bit loads, split it in registers and write the
//     two bf16 tiles straight into the 128B-swizzled K-major layout the MMA
//     descriptor expects;
//   * one elected thread issues tcgen05.mma (M=128, N=BN, K=16) x 4 k-steps x 3
//     products per 64-wide k-block; tcgen05.commit releases the smem stage and,
//     after the last k-block, publishes the accumulator;
//   * 4 warps read the accumulator back with tcgen05.ld (32 lanes x 32 columns
//     per instruction) and run the same fused epilogue as the SGEMM.
//
// Warp roles (320 threads): warp 0 = TMA producer, warp 1 = TMEM alloc + MMA
// issuer, warps 2..9 = A converters, warps 2..5 = epilogue.
#include "kernels.cuh"
#include "gemm_epi.cuh"
#include <cuda.h>
#include <cuda_bf16.h>

namespace ptts {

namespace {

constexpr int TC_BM = 128;
constexpr int TC_BK = 64; // bf16 elements per k-block = one 128-byte swizzle row
constexpr int TC_CONV_THREADS = 256;
constexpr int TC_THREADS = 64 + TC_CONV_THREADS;

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
// x = hi + lo with hi = bf16_rn(x), lo = bf16_rn(x - hi).  Packed conversions
// (cvt.rn.bf16x2.f32 -> one F2FP per pair) instead of scalar F2F: the scalar
// form ran on the slow conversion pipe and bound the whole k-block.
__device__ __forceinline__ void split2(float x0, float x1, uint32_t &hi, uint32_t &lo) {
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(x1), "f"(x0)); // x0 -> low half
  const float h0 = __uint_as_float(hi << 16), h1 = __uint_as_float(hi & 0xFFFF0000u);
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(x1 - h1), "f"(x0 - h0));
}
__device__ __forceinline__ void split8(const float4 &u, const float4 &v, uint4 &hi, uint4 &lo) {
  split2(u.x, u.y, hi.x, lo.x);
  split2(u.z, u.w, hi.y, lo.y);
  split2(v.x, v.y, hi.z, lo.z);
  split2(v.z, v.w, hi.w, lo.w);
}

template <int BN, int STAGES>
struct TcSmem {
  static constexpr int A_TILE = TC_BM * TC_BK * 2; // bytes per bf16 A tile (16 KB)
  static constexpr int W_TILE = BN * TC_BK * 2;
  static constexpr int STAGE = 2 * A_TILE + 2 * W_TILE;
  static constexpr int BYTES = STAGES * STAGE + 1024 /*align*/ + 256 /*barriers*/;
};

template <int BN, int STAGES>
__global__ void __launch_bounds__(TC_THREADS, 1)
    tc_gemm_kernel(const __grid_constant__ CUtensorMap wmap, GemmArgs g, int n_rows_hi) {
  using SM = TcSmem<BN, STAGES>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t *full = reinterpret_cast<uint64_t *>(smem + STAGES * SM::STAGE);
  uint64_t *empty = full + STAGES;
  uint64_t *acc_full = empty + STAGES;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(acc_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.y * TC_BM, n0 = blockIdx.x * BN;
  const int M = g.M, N = g.N, K = g.K;
  const int nkb_all = (K + TC_BK - 1) / TC_BK;
  const int S = gridDim.z;
  const int kb_per = (nkb_all + S - 1) / S;
  const int kb_begin = blockIdx.z * kb_per;
  const int nkb = max(0, min(nkb_all - kb_begin, kb_per)); // k-blocks of this split (host keeps it > 0)
  constexpr uint32_t TMEM_COLS = BN < 32 ? 32 : BN;
  __shared__ int s_last;

  if (threadIdx.x == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&wmap) : "memory");
    for (int s = 0; s < STAGES; s++) {
      mbar_init(&full[s], TC_CONV_THREADS + 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(acc_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== TMA producer: W hi + lo tiles of this k-block =====
    if (lane == 0) {
      for (int kb = 0; kb < nkb; kb++) {
        const int s = kb % STAGES;
        if (kb >= STAGES)
          mbar_wait(&empty[s], ((kb / STAGES) - 1) & 1);
        uint8_t *st = smem + s * SM::STAGE;
        mbar_arrive_expect_tx(&full[s], 2 * SM::W_TILE);
        tma_load_2d(st + 2 * SM::A_TILE, &wmap, &full[s], (kb_begin + kb) * TC_BK, n0);
        tma_load_2d(st + 2 * SM::A_TILE + SM::W_TILE, &wmap, &full[s], (kb_begin + kb) * TC_BK, n_rows_hi + n0);
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(BN);
      for (int kb = 0; kb < nkb; kb++) {
        const int s = kb % STAGES;
        mbar_wait(&full[s], (kb / STAGES) & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t a_hi = smem_u32(smem + s * SM::STAGE);
        const uint32_t a_lo = a_hi + SM::A_TILE;
        const uint32_t w_hi = a_hi + 2 * SM::A_TILE;
        const uint32_t w_lo = w_hi + SM::W_TILE;
#pragma unroll
        for (int k = 0; k < TC_BK / 16; k++) {
          const uint32_t ko = k * 32; // 16 bf16 = 32 bytes along the swizzled row
          const uint64_t dah = make_smem_desc(a_hi + ko), dal = make_smem_desc(a_lo + ko);
          const uint64_t dwh = make_smem_desc(w_hi + ko), dwl = make_smem_desc(w_lo + ko);
          umma_bf16(tmem_base, dal, dwh, idesc, (kb | k) != 0);
          umma_bf16(tmem_base, dah, dwl, idesc, 1);
          umma_bf16(tmem_base, dah, dwh, idesc, 1);
        }
        umma_commit(&empty[s]); // frees the smem stage once these MMAs have read it
      }
      umma_commit(acc_full);
    }
  } else {
    // ===== A converters: fp32 global -> (hi, lo) bf16 swizzled smem =====
    const int ct = threadIdx.x - 64;
    // Three k-blocks of A stay in flight per thread (3 x 8 float4 registers): a
    // single outstanding block left the loop bound by L2/HBM latency.
    constexpr int PF = 3;
    float4 r[PF][4][2];
    auto load_kb = [&](int kb, float4(&dst)[4][2]) {
#pragma unroll
      for (int p = 0; p < 4; p++) {
        const int idx = p * TC_CONV_THREADS + ct;
        const int row = idx >> 3, c = idx & 7;
        const int gm = m0 + row, gk = (kb_begin + kb) * TC_BK + c * 8;
        float4 u = make_float4(0.f, 0.f, 0.f, 0.f), v = u;
        if (gm < M) {
          const float *src = g.A + (long)gm * g.lda + gk;
          if (gk < K)
            u = *reinterpret_cast<const float4 *>(src);
          if (gk + 4 < K)
            v = *reinterpret_cast<const float4 *>(src + 4);
        }
        dst[p][0] = u;
        dst[p][1] = v;
      }
    };
    auto convert_kb = [&](int kb, float4(&src)[4][2]) {
      const int s = kb % STAGES;
      if (kb >= STAGES)
        mbar_wait(&empty[s], ((kb / STAGES) - 1) & 1);
      uint8_t *a_hi = smem + s * SM::STAGE, *a_lo = a_hi + SM::A_TILE;
#pragma unroll
      for (int p = 0; p < 4; p++) {
        const int idx = p * TC_CONV_THREADS + ct;
        const int row = idx >> 3, c = idx & 7;
        uint4 hi, lo;
        split8(src[p][0], src[p][1], hi, lo);
        const int off = (row >> 3) * 1024 + (row & 7) * 128 + ((c ^ (row & 7)) << 4);
        *reinterpret_cast<uint4 *>(a_hi + off) = hi;
        *reinterpret_cast<uint4 *>(a_lo + off) = lo;
      }
      if (kb + PF < nkb)
        load_kb(kb + PF, src);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); // generic writes -> tensor-core reads
      mbar_arrive(&full[s]);
    };
#pragma unroll
    for (int j = 0; j < PF; j++)
      if (j < nkb)
        load_kb(j, r[j]);
    for (int kb0 = 0; kb0 < nkb; kb0 += PF) {
#pragma unroll
      for (int j = 0; j < PF; j++)
        if (kb0 + j < nkb)
          convert_kb(kb0 + j, r[j]);
    }
    // ===== epilogue: TMEM -> smem staging (warps 2..5), then all 8 warps run
    //       the coalesced, vectorised epilogue -> HBM =====
    constexpr int LDS = BN + 4; // padded staging row stride (floats)
    if (warp < 6) {
      mbar_wait(acc_full, 0);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      // One accumulator row per thread -> padded smem tile.  (Writing C straight
      // from the row-per-thread registers scattered every store over 32 rows.)
      // The operand stages are free once acc_full fired.
      const int q = warp & 3; // TMEM lane quarter this warp may access
      float *stg = reinterpret_cast<float *>(smem) + (size_t)q * 32 * LDS;
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 32) {
        uint32_t v[32];
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0;
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
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          *reinterpret_cast<uint4 *>(stg + (size_t)lane * LDS + c0 + j) = make_uint4(v[j], v[j + 1], v[j + 2], v[j + 3]);
      }
    }
    asm volatile("bar.sync 1, 256;" ::: "memory"); // staging tile complete (warps 2..9)
    {
      const GemmEpi &e = g.epi;
      const float *stile = reinterpret_cast<const float *>(smem);
      const int w8 = warp - 2; // 0..7: 16 tile rows each
      // coalesced phase: a warp instruction covers RPI rows x BN columns
      constexpr int F4 = BN / 4, RPI = 32 / F4;
      const int rsub = lane / F4, cf = (lane % F4) * 4;
      const int n = n0 + cf;
      const int tile = blockIdx.y * gridDim.x + blockIdx.x;
      float *tile_ws = g.ws + ((long)tile * S) * TC_BM * BN;
      bool do_epilogue = true;
      if (S > 1) {
        // split-K: publish the partial tile; the last CTA to arrive sums the
        // partials in split order (deterministic) and runs the epilogue.
        float *mine = tile_ws + (long)blockIdx.z * TC_BM * BN;
#pragma unroll 4
        for (int r0 = 0; r0 < 16; r0 += RPI) {
          const int row = w8 * 16 + r0 + rsub;
          *reinterpret_cast<float4 *>(mine + (long)row * BN + cf) =
              *reinterpret_cast<const float4 *>(stile + (size_t)row * LDS + cf);
        }
        __threadfence();
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (threadIdx.x == 64) {
          const int prev = atomicAdd(g.counters + tile, 1);
          s_last = (prev == S - 1);
          if (s_last)
            g.counters[tile] = 0;
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
        do_epilogue = s_last != 0;
        if (do_epilogue)
          __threadfence();
      }
      if (do_epilogue) {
#pragma unroll 2
        for (int r0 = 0; r0 < 16; r0 += RPI) {
          const int row = w8 * 16 + r0 + rsub;
          const int m = m0 + row;
          float4 a;
          if (S == 1) {
            a = *reinterpret_cast<const float4 *>(stile + (size_t)row * LDS + cf);
          } else {
            a = make_float4(0.f, 0.f, 0.f, 0.f);
            const float *src = tile_ws + (long)row * BN + cf;
#pragma unroll 4
            for (int sp = 0; sp < S; sp++) {
              const float4 u = __ldcg(reinterpret_cast<const float4 *>(src + (long)sp * TC_BM * BN));
              a.x += u.x; a.y += u.y; a.z += u.z; a.w += u.w;
            }
          }
          if (m < M && n < N) {
            float4 o = epi_vec4(e, a, m, n);
            if (e.rowmask && !e.rowmask[m / e.mask_div])
              o = make_float4(0.f, 0.f, 0.f, 0.f);
            if (e.C)
              *reinterpret_cast<float4 *>(e.C + (long)m * e.ldc + n) = o;
            if (e.C2) {
              float4 u;
              u.x = act_apply(o.x, ACT_ELU);
              u.y = act_apply(o.y, ACT_ELU);
              u.z = act_apply(o.z, ACT_ELU);
              u.w = act_apply(o.w, ACT_ELU);
              *reinterpret_cast<float4 *>(e.C2 + (long)m * e.ldc2 + n) = u;
            }
          }
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

template <int BN, int STAGES>
int launch_tc(const TcWeight &w, const GemmArgs &g, cudaStream_t st) {
  using SM = TcSmem<BN, STAGES>;
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(tc_gemm_kernel<BN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM::BYTES) !=
        cudaSuccess)
      return -1;
    configured = true;
  }
  const CUtensorMap *map = reinterpret_cast<const CUtensorMap *>(BN == 128 ? w.map128 : (BN == 64 ? w.map64 : w.map32));
  dim3 grid((g.N + BN - 1) / BN, (g.M + TC_BM - 1) / TC_BM, 1);
  // split-K when the tile grid cannot keep the SMs streaming weights
  const long tiles = (long)grid.x * grid.y;
  const int nkb = (g.K + TC_BK - 1) / TC_BK;
  int S = 1;
  if (g.ws && g.counters && tiles < 100 && tiles <= g.n_counters && nkb >= 4) {
    S = (int)((148 + tiles - 1) / tiles);
    if (S > nkb / 2) S = nkb / 2;
    if (S > 4) S = 4;
    while (S > 1 && tiles * S * (long)TC_BM * BN > g.ws_floats) S--;
    while (S > 1 && (long)((nkb + S - 1) / S) * (S - 1) >= nkb) S--; // every split owns >= 1 k-block
    if (S < 1) S = 1;
  }
  grid.z = S;
  tc_gemm_kernel<BN, STAGES><<<grid, TC_THREADS, SM::BYTES, st>>>(*map, g, w.N);
  return 1;
}

} // namespace

// Builds the three TMA descriptors (box 64 x {32,64,128}) over the split weight
// matrix [2N, K] bf16.  Returns 0 on success.
int tc_weight_init(TcWeight *w, const void *d_split, int N, int K) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc)
    return -1;
  w->d_split = d_split;
  w->N = N;
  w->K = K;
  const int boxes[3] = {32, 64, 128};
  void *maps[3] = {w->map32, w->map64, w->map128};
  for (int i = 0; i < 3; i++) {
    cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)(2 * (cuuint64_t)N)};
    cuuint64_t strides[1] = {(cuuint64_t)K * 2};
    cuuint32_t box[2] = {(cuuint32_t)TC_BK, (cuuint32_t)boxes[i]};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(reinterpret_cast<CUtensorMap *>(maps[i]), CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2,
                     const_cast<void *>(d_split), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS)
      return -2;
  }
  return 0;
}

bool tc_gemm_supported(const GemmArgs &g) {
  return g.M >= 1 && g.N >= 32 && g.N % 16 == 0 && g.K >= 32 && g.K % 8 == 0 && g.lda % 4 == 0 &&
         (reinterpret_cast<uintptr_t>(g.A) & 15) == 0;
}

int tc_gemm(const TcWeight &w, const GemmArgs &g, cudaStream_t st, const char **variant) {
  // Narrower N tiles first (more CTAs streaming weights when M is skinny), then a
  // small split-K; wide tiles when the grid is already large (less A re-reading).
  const long mt = (g.M + TC_BM - 1) / TC_BM;
  auto tiles = [&](int bn) { return mt * ((g.N + bn - 1) / bn); };
  int bn = 128;
  if (g.N < 128 || tiles(128) < 96)
    bn = 64;
  if (g.N < 64 || (bn == 64 && tiles(64) < 96))
    bn = 32;
  if (bn == 128) {
    *variant = "tcgen05_128x128";
    return launch_tc<128, 3>(w, g, st);
  }
  if (bn == 64) {
    *variant = "tcgen05_128x64";
    return launch_tc<64, 4>(w, g, st);
  }
  *variant = "tcgen05_128x32";
  return launch_tc<32, 4>(w, g, st);
}

} // namespace ptts
