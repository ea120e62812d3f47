// resblock.cu — one SEANet residual block as ONE persistent tcgen05 kernel
// (internal/native/mimi.go:146-164:  x + conv2(ELU(conv1(ELU(x)))), conv1 k=3 C -> C/2, conv2
// k=1 C/2 -> C, followed by the ELU that feeds the next layer):
//
//   out = elu( y + b2 + W2 . elu( b1 + W1 . [e(r-2), e(r-1), e(r)] ) ),   e = elu(y)
//
// Both contractions run on the tensor cores with split-bf16 operands (3 products each, as
// tcgemm.cu) and the hidden activation never leaves the SM: GEMM1's accumulator (TMEM) is
// read by the epilogue warps, biased, ELU'd, split and written as GEMM2's A tile straight
// into swizzled shared memory; GEMM2's accumulator gets the residual + ELU epilogue.
// Unfused, the hidden tensor costs a write + a read of rows*C/2*4 B and a kernel whose
// epilogue is the bottleneck (K = C/2 <= 64); fused, HBM sees e, y and the output once.
//
// Per CTA (192 threads): warp 0 TMA (e tiles through the 3-D overlapping-row map + W1 tiles;
// W2 once), warp 1 MMA issue, warps 2..5 epilogue.  Items (128-row tiles) are pipelined:
// GEMM1(i) is issued before GEMM2(i-1); E1(i) (hidden -> smem) runs before E2(i-1) (output),
// all accumulators and the hidden tile are double-buffered.
#include "kernels.cuh"
#include "launch.cuh"
#include "tc_common.cuh"
#include <stdlib.h>

namespace ptts {

namespace {

// warp 0 TMA, warp 1 MMA, warps 2..5 E1 (hidden tile), warps 6..13 E2 (output; two warps per
// TMEM lane quarter, even / odd 32-column chunks): the two epilogues run concurrently.
#ifndef RB_E2W_64
#define RB_E2W_64 1 // C = 64: 8 E2 warps (1) or 4 (0)
#endif
__host__ __device__ constexpr int rb_e2_warps(int c) { return (RB_E2W_64 || c != 64) ? 8 : 4; }
__host__ __device__ constexpr int rb_threads(int c) { return 64 + 128 + rb_e2_warps(c) * 32; }

template <int C, int H>
struct RbSmem {
  static constexpr int STAGES = C == 64 ? 3 : 2;
  static constexpr int A_TILE = TC_BM * TC_BK * 2;      // 16 KB
  static constexpr int W1_TILE = H * TC_BK * 2;
  static constexpr int STAGE = 2 * A_TILE + 2 * W1_TILE;
  static constexpr int W2_TILE = C * TC_BK * 2;         // [C rows x 64 k] bf16 (k >= H is zero-filled by TMA)
  static constexpr int A2_BUF = 2 * A_TILE;             // hidden tile hi + lo
  static constexpr int STG_LD = 36;
  static constexpr int STG_WARP = 32 * STG_LD * 4;
  static constexpr int OFF_W2 = STAGES * STAGE;
  static constexpr int OFF_A2 = OFF_W2 + 2 * W2_TILE;
  static constexpr int OFF_STG = OFF_A2 + A2_BUF; // one hidden tile: GEMM2 is a handful of MMAs
  static constexpr int OFF_BAR = OFF_STG + rb_e2_warps(C) * STG_WARP;
  static constexpr int BYTES = OFF_BAR + 256 + 1024;
};

struct RbArgs {
  int M, taps, a_lo_row;
  const float *b1, *b2;
  const float *y; // residual input [M, C] fp32
  const uint8_t *rowmask;
  int mask_div;
  float *out;     // elu(..) fp32 [M, C] (optional)
  SplitOut outs;  // its split planes (optional)
};

template <int C, int H>
__global__ void __launch_bounds__(rb_threads(C), 1)
    resblock_kernel(const __grid_constant__ CUtensorMap w1map, const __grid_constant__ CUtensorMap w2map,
                    const __grid_constant__ CUtensorMap amap, RbArgs g) {
  using SM = RbSmem<C, H>;
  constexpr int STAGES = SM::STAGES;
  constexpr uint32_t TMEM_COLS = (2 * H + 2 * C) <= 256 ? 256 : 512;
  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t *full = reinterpret_cast<uint64_t *>(smem + SM::OFF_BAR);
  uint64_t *empty = full + STAGES;
  uint64_t *acc1_full = empty + STAGES; // [2] each below
  uint64_t *acc1_empty = acc1_full + 2;
  uint64_t *a2_full = acc1_empty + 2;
  uint64_t *a2_empty = a2_full + 2;
  uint64_t *acc2_full = a2_empty + 2;
  uint64_t *acc2_empty = acc2_full + 2;
  uint64_t *w2_full = acc2_empty + 2;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(w2_full + 1);

  __shared__ float s_b1[H];
  if (threadIdx.x < H)
    s_b1[threadIdx.x] = g.b1[threadIdx.x];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int M = g.M;
  const int n_items = (M + TC_BM - 1) / TC_BM;
  const int nkb1 = g.taps * C / TC_BK;
  const int my_items = n_items > (int)blockIdx.x ? (n_items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;

  if (threadIdx.x == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&w1map) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&w2map) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&amap) : "memory");
    for (int s = 0; s < STAGES; s++) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int b = 0; b < 2; b++) {
      mbar_init(&acc1_full[b], 1);
      mbar_init(&acc1_empty[b], 128);
      mbar_init(&a2_full[b], 128);
      mbar_init(&a2_empty[b], 1);
      mbar_init(&acc2_full[b], 1);
      mbar_init(&acc2_empty[b], rb_e2_warps(C) * 32);
    }
    mbar_init(w2_full, 1);
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
  const uint32_t tm_acc1 = tmem_base;           // + b * H
  const uint32_t tm_acc2 = tmem_base + 2 * H;   // + b * C

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      mbar_arrive_expect_tx(w2_full, 2 * SM::W2_TILE);
      tma_load_2d(smem + SM::OFF_W2, &w2map, w2_full, 0, 0);
      tma_load_2d(smem + SM::OFF_W2 + SM::W2_TILE, &w2map, w2_full, 0, C);
      uint32_t it = 0;
      for (int k = 0; k < my_items; k++) {
        const int m0 = ((int)blockIdx.x + k * (int)gridDim.x) * TC_BM;
        for (int kb = 0; kb < nkb1; kb++, it++) {
          const int s = (int)(it % STAGES);
          if (it >= STAGES)
            mbar_wait(&empty[s], (uint32_t)((it / STAGES) - 1) & 1);
          uint8_t *st = smem + s * SM::STAGE;
          mbar_arrive_expect_tx(&full[s], SM::STAGE);
          const int k0 = kb * TC_BK;
          const int tap = k0 / C, c0 = k0 - tap * C;
          tma_load_3d(st, &amap, &full[s], c0, tap, m0);
          tma_load_3d(st + SM::A_TILE, &amap, &full[s], c0, tap, g.a_lo_row + m0);
          tma_load_2d(st + 2 * SM::A_TILE, &w1map, &full[s], k0, 0);
          tma_load_2d(st + 2 * SM::A_TILE + SM::W1_TILE, &w1map, &full[s], k0, H);
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: GEMM1(i), then GEMM2(i-1) =====
    if (lane == 0) {
      constexpr uint32_t idesc1 = make_idesc(H), idesc2 = make_idesc(C);
      auto gemm2 = [&](int j) {
        const int b = j & 1;
        mbar_wait(&a2_full[0], (uint32_t)j & 1);
        if (j >= 2)
          mbar_wait(&acc2_empty[b], (uint32_t)((j >> 1) - 1) & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t a_hi = smem_u32(smem + SM::OFF_A2), a_lo = a_hi + SM::A_TILE;
        const uint32_t w_hi = smem_u32(smem + SM::OFF_W2), w_lo = w_hi + SM::W2_TILE;
#pragma unroll
        for (int k = 0; k < H / 16; k++) {
          const uint32_t ko = k * 32;
          const uint64_t dah = make_smem_desc(a_hi + ko), dal = make_smem_desc(a_lo + ko);
          const uint64_t dwh = make_smem_desc(w_hi + ko), dwl = make_smem_desc(w_lo + ko);
          umma_bf16(tm_acc2 + b * C, dal, dwh, idesc2, k != 0);
          umma_bf16(tm_acc2 + b * C, dah, dwl, idesc2, 1);
          umma_bf16(tm_acc2 + b * C, dah, dwh, idesc2, 1);
        }
        umma_commit(&acc2_full[b]);
        umma_commit(&a2_empty[0]); // the hidden tile may be rewritten once these MMAs have read it
      };
      mbar_wait(w2_full, 0);
      uint32_t it = 0;
      for (int k = 0; k < my_items; k++) {
        const int b = k & 1;
        if (k >= 2) {
          mbar_wait(&acc1_empty[b], (uint32_t)((k >> 1) - 1) & 1);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        }
        for (int kb = 0; kb < nkb1; kb++, it++) {
          const int s = (int)(it % STAGES);
          mbar_wait(&full[s], (uint32_t)(it / STAGES) & 1);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t a_hi = smem_u32(smem + s * SM::STAGE), a_lo = a_hi + SM::A_TILE;
          const uint32_t w_hi = a_hi + 2 * SM::A_TILE, w_lo = w_hi + SM::W1_TILE;
#pragma unroll
          for (int kk = 0; kk < TC_BK / 16; kk++) {
            const uint32_t ko = kk * 32;
            const uint64_t dah = make_smem_desc(a_hi + ko), dal = make_smem_desc(a_lo + ko);
            const uint64_t dwh = make_smem_desc(w_hi + ko), dwl = make_smem_desc(w_lo + ko);
            umma_bf16(tm_acc1 + b * H, dal, dwh, idesc1, (kb | kk) != 0);
            umma_bf16(tm_acc1 + b * H, dah, dwl, idesc1, 1);
            umma_bf16(tm_acc1 + b * H, dah, dwh, idesc1, 1);
          }
          umma_commit(&empty[s]);
        }
        umma_commit(&acc1_full[b]);
        if (k >= 1)
          gemm2(k - 1);
      }
      if (my_items >= 1)
        gemm2(my_items - 1);
    }
  } else {
    // ===== epilogue warps: E1(i) hidden -> smem, then E2(i-1) output =====
    const int q = warp & 3;
    const int row = q * 32 + lane; // TMEM lane = tile row this thread owns in E1
    const bool is_e2 = warp >= 6;
    const uint32_t stg_addr = smem_u32(smem + SM::OFF_STG + (is_e2 ? warp - 6 : 0) * SM::STG_WARP);
    const int c_first = is_e2 ? ((warp - 6) >> 2) * 32 : 0;
    constexpr int C_STEP = rb_e2_warps(C) == 8 ? 64 : 32;
    const int rsub = lane >> 3, cf = (lane & 7) * 4;
    GemmEpi e; // epilogue-2 description for the shared row routines
    e.C2 = g.out;
    e.ldc2 = C;
    e.C2s = g.outs;
    const int epi_mask = (g.out ? 4 : 0) | (g.outs.hi ? 8 : 0);
    auto e1 = [&](int k) {
      const int b = k & 1;
      mbar_wait(&acc1_full[b], (uint32_t)(k >> 1) & 1);
      if (k >= 1)
        mbar_wait(&a2_empty[0], (uint32_t)(k - 1) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      uint8_t *a_hi = smem + SM::OFF_A2, *a_lo = a_hi + SM::A_TILE;
#pragma unroll 1
      for (int c0 = 0; c0 < H; c0 += 32) {
        uint32_t v[32];
        const uint32_t taddr = tm_acc1 + b * H + ((uint32_t)(q * 32) << 16) + (uint32_t)c0;
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
        // this thread's row, 32 hidden channels: bias, ELU, split, K-major 128B-swizzled store
#pragma unroll
        for (int j = 0; j < 32; j += 8) {
          float x[8];
#pragma unroll
          for (int t = 0; t < 8; t++)
            x[t] = elu_fast(__uint_as_float(v[j + t]) + s_b1[c0 + j + t]);
          uint4 hi, lo;
          split2(x[0], x[1], hi.x, lo.x);
          split2(x[2], x[3], hi.y, lo.y);
          split2(x[4], x[5], hi.z, lo.z);
          split2(x[6], x[7], hi.w, lo.w);
          const int chunk = (c0 + j) >> 3; // 16-byte chunk inside the 128-byte row
          const int off = (row >> 3) * 1024 + (row & 7) * 128 + ((chunk ^ (row & 7)) << 4);
          *reinterpret_cast<uint4 *>(a_hi + off) = hi;
          *reinterpret_cast<uint4 *>(a_lo + off) = lo;
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      mbar_arrive(&acc1_empty[b]);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); // generic writes -> tensor-core reads
      mbar_arrive(&a2_full[0]);
    };
    auto e2 = [&](int k) {
      const int b = k & 1;
      const int m0 = ((int)blockIdx.x + k * (int)gridDim.x) * TC_BM;
      uint32_t rows = 0;
      // one integer division per tile and thread instead of eight (the epilogue is what bounds this kernel):
      // the eight rows are 4 apart, so quotient and remainder by mask_div are carried along
      int mq = 0, mr = 0;
      if (g.rowmask) {
        const int mb = m0 + q * 32 + rsub;
        mq = mb / g.mask_div;
        mr = mb - mq * g.mask_div;
      }
#pragma unroll
      for (int j = 0; j < 8; j++) {
        const int m = m0 + q * 32 + j * 4 + rsub;
        if (m < M) {
          rows |= 1u << j;
          if (!g.rowmask || g.rowmask[mq])
            rows |= 0x100u << j;
        }
        mr += 4;
        while (mr >= g.mask_div) {
          mr -= g.mask_div;
          mq++;
        }
      }
      const bool fullrows = __all_sync(0xffffffffu, (rows & 0xFFu) == 0xFFu);
      mbar_wait(&acc2_full[b], (uint32_t)(k >> 1) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
      for (int c0 = c_first; c0 < C; c0 += C_STEP) {
        const int n = c0 + cf;
        uint32_t v[32];
        const uint32_t taddr = tm_acc2 + b * C + ((uint32_t)(q * 32) << 16) + (uint32_t)c0;
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
        float4 res[8];
#pragma unroll
        for (int j = 0; j < 8; j++) {
          const int m = m0 + q * 32 + j * 4 + rsub;
          res[j] = ((rows >> j) & 1) ? *reinterpret_cast<const float4 *>(g.y + (long)m * C + n)
                                     : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        const float4 bias4 = __ldg(reinterpret_cast<const float4 *>(g.b2 + n));
        const float4 cs4 = make_float4(1.f, 1.f, 1.f, 1.f);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          sts128(stg_addr + (uint32_t)(lane * SM::STG_LD + j) * 4, v[j], v[j + 1], v[j + 2], v[j + 3]);
        __syncwarp();
        const uint32_t rd_addr = stg_addr + (uint32_t)(rsub * SM::STG_LD + cf) * 4;
        const long m_first = m0 + q * 32 + rsub;
        if (epi_mask == 4) {
          if (fullrows)
            epi_rows_fast<ACT_NONE, 4, true>(e, rd_addr, SM::STG_LD * 4, m_first, n, rows, bias4, cs4, res, true);
          else
            epi_rows_fast<ACT_NONE, 4, false>(e, rd_addr, SM::STG_LD * 4, m_first, n, rows, bias4, cs4, res, true);
        } else if (epi_mask == 8) {
          if (fullrows)
            epi_rows_fast<ACT_NONE, 8, true>(e, rd_addr, SM::STG_LD * 4, m_first, n, rows, bias4, cs4, res, true);
          else
            epi_rows_fast<ACT_NONE, 8, false>(e, rd_addr, SM::STG_LD * 4, m_first, n, rows, bias4, cs4, res, true);
        } else {
          epi_rows_fast<ACT_NONE, 12, false>(e, rd_addr, SM::STG_LD * 4, m_first, n, rows, bias4, cs4, res, true);
        }
        __syncwarp();
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      mbar_arrive(&acc2_empty[b]);
    };
    if (!is_e2) {
      for (int k = 0; k < my_items; k++)
        e1(k);
    } else {
      for (int k = 0; k < my_items; k++)
        e2(k);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

template <int C, int H>
int launch_rb(const TcWeight &w1, const TcWeight &w2, const void *amap, const RbArgs &a, cudaStream_t st) {
  using SM = RbSmem<C, H>;
  static volatile bool configured[64] = {};
  volatile bool &conf = device_flag(configured);
  if (!conf) {
    if (cudaFuncSetAttribute(resblock_kernel<C, H>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM::BYTES) != cudaSuccess)
      return -1;
    conf = true;
  }
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int items = (a.M + TC_BM - 1) / TC_BM;
  const int grid = items < sms ? items : sms;
  const CUtensorMap *m1 = reinterpret_cast<const CUtensorMap *>(H == 32 ? w1.map32 : w1.map64);
  const CUtensorMap *m2 = reinterpret_cast<const CUtensorMap *>(C == 64 ? w2.map64 : w2.map128);
  launch_k(resblock_kernel<C, H>, dim3(grid), dim3(rb_threads(C)), SM::BYTES, st, *m1, *m2, *reinterpret_cast<const CUtensorMap *>(amap),
           a);
  return 1;
}

} // namespace

// Fused residual block for C in {64, 128} (hidden C/2, conv1 with 3 taps, conv2 k=1).
// w1: [C/2, 3*C] split weight, w2: [C, C/2]; amap: the {C, 3, rows} conv map over the split
// planes of e = elu(y).  Returns 0 without launching for other shapes.
int resblock_fused(const TcWeight &w1, const TcWeight &w2, const void *amap, int a_lo_row, int M, int C, int H, int taps,
                   const float *b1, const float *b2, const float *y, const uint8_t *rowmask, int mask_div, float *out,
                   SplitOut outs, cudaStream_t st) {
  if (!amap || !b1 || !b2 || taps != 3 || M <= 0 || H * 2 != C || (C != 64 && C != 128) || w1.K != taps * C || w1.N != H ||
      w2.K != H || w2.N != C)
    return 0;
  RbArgs a;
  a.M = M; a.taps = taps; a.a_lo_row = a_lo_row;
  a.b1 = b1; a.b2 = b2; a.y = y;
  a.rowmask = rowmask; a.mask_div = mask_div;
  a.out = out; a.outs = outs;
  const int n = C == 64 ? launch_rb<64, 32>(w1, w2, amap, a, st) : launch_rb<128, 64>(w1, w2, amap, a, st);
  return n < 0 ? 0 : n;
}

} // namespace ptts
