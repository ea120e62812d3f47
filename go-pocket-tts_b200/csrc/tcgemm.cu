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
#include "split.cuh"
#include "launch.cuh"
#include "tc_common.cuh"
#include <cuda.h>
#include <cuda_bf16.h>
#include <stdio.h>
#include <stdlib.h>
#include <type_traits>

namespace ptts {

namespace {

int tc_num_sms() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
      n = 148;
  }
  return n;
}

__device__ __forceinline__ unsigned long long gtimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define TC_STAMP(i)                                                            \
  do {                                                                         \
    if (g.dbg && blockIdx.x == 0)                                              \
      g.dbg[i] = gtimer();                                                     \
  } while (0)

template <int BN, int STAGES, int EW = 4>
struct TcSmem {
  static constexpr int A_TILE = TC_BM * TC_BK * 2; // bytes per bf16 A tile (16 KB)
  static constexpr int W_TILE = BN * TC_BK * 2;
  static constexpr int STAGE = 2 * A_TILE + 2 * W_TILE;
  static constexpr int STG_LD = 36;                 // staging row stride (floats): conflict-free v4 both ways
  static constexpr int STG_WARP = 32 * STG_LD * 4;  // one 32x32 fp32 chunk per epilogue warp
  static constexpr int BYTES = STAGES * STAGE + EW * STG_WARP + 1024 /*align*/ + 256 /*barriers*/;
};

// Work decomposition shared by every warp role: item i -> (m tile, n tile, k split).
// Splits of one tile are adjacent items so they run at the same time.
struct TcItems {
  int total;
  int ntn, S, kb_per, nkb_all;
  int epi_kind; // epi_code() of a compiled epilogue, or -1: generic
  int cluster;  // 1: the S splits of a tile are one thread-block cluster and reduce through DSMEM
  __device__ __forceinline__ void decode(int i, int &m0, int &n0, int &kb_begin, int &nkb, int &z, int &tile,
                                         int bn) const {
    z = (int)(i % S);
    tile = i / S;
    const int nt = (int)(tile % ntn);
    const int mt = (int)(tile / ntn);
    m0 = mt * TC_BM;
    n0 = nt * bn;
    kb_begin = z * kb_per;
    nkb = min(nkb_all - kb_begin, kb_per);
  }
};

// Persistent, warp-specialised kernel: each CTA walks items blockIdx.x, +gridDim.x, ...
// The accumulator is double-buffered in TMEM so the epilogue of item i overlaps the
// main loop of item i+1; the smem operand ring runs straight through item boundaries.
// ATMA: A is already split into bf16 planes behind `amap`; the TMA warp fetches both
// operands and the converter warps are not launched.
// WLO = false (own instantiations, TMA-fed only): the weight's lo plane is identically zero and is neither
// fetched nor multiplied — Ahi*Wlo would add exact zeros, so the result is bit-identical to the 3-product form.
template <int BN, int STAGES, bool ATMA, int EW, bool WLO = true>
__global__ void __launch_bounds__(tc_threads(EW, ATMA), 1)
    tc_gemm_kernel(const __grid_constant__ CUtensorMap wmap, const __grid_constant__ CUtensorMap amap, GemmArgs g,
                   int n_rows_hi, TcItems items) {
  using SM = TcSmem<BN, STAGES, EW>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t *stg_base = smem + STAGES * SM::STAGE;
  uint64_t *full = reinterpret_cast<uint64_t *>(stg_base + EW * SM::STG_WARP);
  uint64_t *empty = full + STAGES;
  uint64_t *acc_full = empty + STAGES; // [2]
  uint64_t *acc_empty = acc_full + 2;  // [2]
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(acc_empty + 2);
  int *s_last = reinterpret_cast<int *>(tmem_slot + 1);

  PDL_TRIGGER(); // the next grid may be scheduled; it blocks in its own PDL_WAIT until this one is done
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int M = g.M, N = g.N, K = g.K;
  const int S = items.S;
  constexpr uint32_t TMEM_COLS = 2 * BN < 32 ? 32 : 2 * BN;

  if (threadIdx.x == 0) {
    TC_STAMP(0);
    asm volatile("prefetch.tensormap [%0];" ::"l"(&wmap) : "memory");
    if (ATMA)
      asm volatile("prefetch.tensormap [%0];" ::"l"(&amap) : "memory");
    for (int s = 0; s < STAGES; s++) {
      mbar_init(&full[s], ATMA ? 1 : TC_CONV_THREADS + 1);
      mbar_init(&empty[s], 1);
    }
    for (int b = 0; b < 2; b++) {
      mbar_init(&acc_full[b], 1);
      mbar_init(&acc_empty[b], EW * 32);
    }
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
  if (threadIdx.x == 0)
    TC_STAMP(1);

  if (warp == 0) {
    // ===== TMA producer: W (and, ATMA, A) hi + lo tiles of every k-block of every item.
    // The weight tiles of the first ring pass do not depend on the previous kernel, so
    // they are requested before PDL_WAIT; their A tiles follow once it returns. =====
    if (lane == 0) {
      uint32_t it = 0;
      bool waited = false;
      int npend = 0;
      int pend_s[STAGES], pend_k0[STAGES], pend_m0[STAGES];
      auto load_a = [&](int s, int k0, int m0) {
        uint8_t *st = smem + s * SM::STAGE;
        if (g.a_cin > 0) { // conv over overlapping rows: k = tap * cin + c
          const int tap = k0 / g.a_cin, c0 = k0 - tap * g.a_cin;
          tma_load_3d(st, &amap, &full[s], c0, tap, m0);
          tma_load_3d(st + SM::A_TILE, &amap, &full[s], c0, tap, g.a_lo_row + m0);
        } else {
          tma_load_2d(st, &amap, &full[s], k0, m0);
          tma_load_2d(st + SM::A_TILE, &amap, &full[s], k0, g.a_lo_row + m0);
        }
      };
      auto dep_wait = [&]() {
        PDL_WAIT();
        waited = true;
        for (int i = 0; i < npend; i++)
          load_a(pend_s[i], pend_k0[i], pend_m0[i]);
        npend = 0;
      };
      for (int i = blockIdx.x; i < items.total; i += gridDim.x) {
        int m0, n0, kb_begin, nkb, z;
        int tile;
        items.decode(i, m0, n0, kb_begin, nkb, z, tile, BN);
        for (int kb = 0; kb < nkb; kb++, it++) {
          const int s = (int)(it % STAGES);
          if (it >= STAGES) {
            if (!waited)
              dep_wait();
            mbar_wait(&empty[s], (uint32_t)((it / STAGES) - 1) & 1);
          }
          uint8_t *st = smem + s * SM::STAGE;
          if (it < 32)
            TC_STAMP(96 + it);
          mbar_arrive_expect_tx(&full[s], (WLO ? 2 : 1) * SM::W_TILE + (ATMA ? 2 * SM::A_TILE : 0));
          const int k0 = (kb_begin + kb) * TC_BK;
          if (ATMA) {
            if (waited)
              load_a(s, k0, m0);
            else {
              pend_s[npend] = s;
              pend_k0[npend] = k0;
              pend_m0[npend] = m0;
              npend++;
            }
          }
          tma_load_2d(st + 2 * SM::A_TILE, &wmap, &full[s], k0, n0);
          if (WLO)
            tma_load_2d(st + 2 * SM::A_TILE + SM::W_TILE, &wmap, &full[s], k0, n_rows_hi + n0);
        }
      }
      if (!waited)
        dep_wait();
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(BN);
      long it = 0, li = 0;
      for (long i = blockIdx.x; i < items.total; i += gridDim.x, li++) {
        int m0, n0, kb_begin, nkb, z;
        int tile;
        items.decode(i, m0, n0, kb_begin, nkb, z, tile, BN);
        const int b = (int)(li & 1);
        if (li >= 2) { // the epilogue must have drained this accumulator buffer
          mbar_wait(&acc_empty[b], (uint32_t)((li >> 1) - 1) & 1);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        }
        const uint32_t tmem_d = tmem_base + (uint32_t)(b * BN);
        for (int kb = 0; kb < nkb; kb++, it++) {
          const int s = (int)(it % STAGES);
          mbar_wait(&full[s], (uint32_t)(it / STAGES) & 1);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          if (it < 32)
            TC_STAMP(2 + it);
          const uint32_t a_hi = smem_u32(smem + s * SM::STAGE);
          const uint32_t a_lo = a_hi + SM::A_TILE;
          const uint32_t w_hi = a_hi + 2 * SM::A_TILE;
          const uint32_t w_lo = w_hi + SM::W_TILE;
#pragma unroll
          for (int k = 0; k < TC_BK / 16; k++) {
            const uint32_t ko = k * 32; // 16 bf16 = 32 bytes along the swizzled row
            const uint64_t dah = make_smem_desc(a_hi + ko), dal = make_smem_desc(a_lo + ko);
            const uint64_t dwh = make_smem_desc(w_hi + ko), dwl = make_smem_desc(w_lo + ko);
            umma_bf16(tmem_d, dal, dwh, idesc, (kb | k) != 0);
            if (WLO)
              umma_bf16(tmem_d, dah, dwl, idesc, 1);
            umma_bf16(tmem_d, dah, dwh, idesc, 1);
          }
          umma_commit(&empty[s]); // frees the smem stage once these MMAs have read it
        }
        umma_commit(&acc_full[b]);
        if (li == 0)
          TC_STAMP(40);
      }
    }
  } else if (warp >= tc_base_threads(EW) / 32) {
    // ===== A converters: fp32 global -> (hi, lo) bf16 swizzled smem =====
    PDL_WAIT();
    const int ct = threadIdx.x - tc_base_threads(EW);
    // PF k-blocks of A stay in flight per thread, across item boundaries.
    constexpr int PF = 3;
    float4 r[PF][4][2];
    struct Cur {
      long item;
      int kb, nkb, m0, kb_begin;
    };
    auto setup = [&](Cur &c) {
      if (c.item < items.total) {
        int n0, z;
        int tile;
        items.decode(c.item, c.m0, n0, c.kb_begin, c.nkb, z, tile, BN);
      } else
        c.nkb = 0;
      c.kb = 0;
    };
    auto advance = [&](Cur &c) {
      if (++c.kb >= c.nkb) {
        c.item += gridDim.x;
        setup(c);
      }
    };
    auto load_kb = [&](const Cur &c, float4(&dst)[4][2]) {
#pragma unroll
      for (int p = 0; p < 4; p++) {
        const int idx = p * TC_CONV_THREADS + ct;
        const int row = idx >> 3, cc = idx & 7;
        const int gm = c.m0 + row, gk = (c.kb_begin + c.kb) * TC_BK + cc * 8;
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
    Cur lc, cc;
    lc.item = blockIdx.x;
    setup(lc);
    cc = lc;
#pragma unroll
    for (int j = 0; j < PF; j++) {
      if (lc.nkb > 0) {
        load_kb(lc, r[j]);
        advance(lc);
      }
    }
    long it = 0;
    while (cc.nkb > 0) {
#pragma unroll
      for (int j = 0; j < PF; j++) {
        if (cc.nkb > 0) {
          const int s = (int)(it % STAGES);
          if (it >= STAGES)
            mbar_wait(&empty[s], (uint32_t)((it / STAGES) - 1) & 1);
          uint8_t *a_hi = smem + s * SM::STAGE, *a_lo = a_hi + SM::A_TILE;
#pragma unroll
          for (int p = 0; p < 4; p++) {
            const int idx = p * TC_CONV_THREADS + ct;
            const int row = idx >> 3, c8 = idx & 7;
            uint4 hi, lo;
            split8(r[j][p][0], r[j][p][1], hi, lo);
            const int off = (row >> 3) * 1024 + (row & 7) * 128 + ((c8 ^ (row & 7)) << 4);
            *reinterpret_cast<uint4 *>(a_hi + off) = hi;
            *reinterpret_cast<uint4 *>(a_lo + off) = lo;
          }
          if (lc.nkb > 0) {
            load_kb(lc, r[j]);
            advance(lc);
          }
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); // generic writes -> tensor-core reads
          if (ct == 0 && it < 32)
            TC_STAMP(48 + it);
          mbar_arrive(&full[s]);
          advance(cc);
          it++;
        }
      }
    }
  } else {
    // ===== epilogue warps: TMEM -> per-warp smem transpose -> coalesced fused epilogue =====
    PDL_WAIT();
    const GemmEpi &e = g.epi;
    const int q = warp & 3; // TMEM lane quarter this warp may access
    const uint32_t stg_addr = smem_u32(stg_base + (warp - 2) * SM::STG_WARP);
    const int et = threadIdx.x - 64; // 0..EW*32-1
    // with 8 epilogue warps, warps 2..5 take the even 32-column chunks, 6..9 the odd ones
    const int c_first = EW == 8 ? ((warp - 2) >> 2) * 32 : 0;
    constexpr int C_STEP = EW == 8 ? 64 : 32;
    const int rsub = lane >> 3, cf = (lane & 7) * 4;     // coalesced phase: 4 rows x 32 columns per instruction
    if (S == 1 && !items.cluster) {
      // ---- software-pipelined epilogue (the common case).  A unit is one 32-column chunk
      // of one item.  The TMEM load of unit u+1 (and its residual rows) is issued right
      // after unit u's accumulator chunk has been parked in the staging buffer, so it is in
      // flight while unit u's rows are processed and stored: the short-K shapes are bound
      // by this warp's latency chain, not by its instruction count. ----
      struct UnitCtx {
        int m0, n0, c0, b;
        uint32_t rows;
        bool full, nok, last_chunk, have_res;
        float4 bias4, cs4;
        float4 res[8];
      };
      const int nch = (BN - c_first + C_STEP - 1) / C_STEP; // chunks of an item this warp owns
      const int my_items = ((int)items.total - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
      const int nunits = c_first < BN ? my_items * nch : 0;
      uint32_t v[32];
      uint32_t rows_item = 0;
      bool full_item = false;
      int im0 = 0, in0 = 0;
      auto issue = [&](int u, UnitCtx &c) {
        const int li2 = u / nch, j = u - li2 * nch;
        c.b = li2 & 1;
        c.c0 = c_first + j * C_STEP;
        c.last_chunk = j == nch - 1;
        if (j == 0) {
          int kb_begin, nkb, z, tile;
          items.decode((int)blockIdx.x + li2 * (int)gridDim.x, im0, in0, kb_begin, nkb, z, tile, BN);
          rows_item = 0;
          // masked (Mimi conv) launches: one integer division per item and thread, quotient / remainder carried
          // along the eight rows (4 apart); unmasked launches (the AR step) skip all of it
          int mq = 0, mr = 0;
          if (e.rowmask) {
            const int mb = im0 + q * 32 + rsub;
            mq = mb / e.mask_div;
            mr = mb - mq * e.mask_div;
          }
#pragma unroll
          for (int jj = 0; jj < 8; jj++) {
            const int m = im0 + q * 32 + jj * 4 + rsub;
            if (m < M) {
              rows_item |= 1u << jj;
              if (!e.rowmask || e.rowmask[mq])
                rows_item |= 0x100u << jj;
            }
            if (e.rowmask) {
              mr += 4;
              while (mr >= e.mask_div) {
                mr -= e.mask_div;
                mq++;
              }
            }
          }
          full_item = __all_sync(0xffffffffu, (rows_item & 0xFFu) == 0xFFu);
          mbar_wait(&acc_full[c.b], (uint32_t)(li2 >> 1) & 1);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        }
        c.m0 = im0;
        c.n0 = in0;
        c.rows = rows_item;
        c.full = full_item;
        const int n = in0 + c.c0 + cf;
        c.nok = n < N;
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(c.b * BN + c.c0);
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
        c.have_res = e.residual != nullptr;
        if (c.have_res) {
#pragma unroll
          for (int jj = 0; jj < 8; jj++) {
            const int m = im0 + q * 32 + jj * 4 + rsub;
            c.res[jj] = (c.nok && ((rows_item >> jj) & 1))
                            ? *reinterpret_cast<const float4 *>(e.residual + (long)m * e.ldr + n)
                            : make_float4(0.f, 0.f, 0.f, 0.f);
          }
        }
        c.bias4 = make_float4(0.f, 0.f, 0.f, 0.f);
        c.cs4 = make_float4(1.f, 1.f, 1.f, 1.f);
        if (c.nok) {
          if (e.bias)
            c.bias4 = __ldg(reinterpret_cast<const float4 *>(e.bias + n));
          if (e.colscale)
            c.cs4 = __ldg(reinterpret_cast<const float4 *>(e.colscale + n));
        }
      };
      UnitCtx cur, nxt;
      if (nunits > 0)
        issue(0, nxt);
      for (int u = 0; u < nunits; u++) {
        cur = nxt;
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          sts128(stg_addr + (uint32_t)(lane * SM::STG_LD + j) * 4, v[j], v[j + 1], v[j + 2], v[j + 3]);
        __syncwarp();
        if (cur.last_chunk) { // this warp has read all of its share of the accumulator
          asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
          mbar_arrive(&acc_empty[cur.b]);
        }
        if (u + 1 < nunits)
          issue(u + 1, nxt);
        if (cur.nok) {
          const uint32_t rd_addr = stg_addr + (uint32_t)(rsub * SM::STG_LD + cf) * 4;
          const long m_first = cur.m0 + q * 32 + rsub;
          epi_rows_dispatch(items.epi_kind, e, rd_addr, SM::STG_LD * 4, m_first, cur.n0 + cur.c0 + cf, cur.rows, cur.bias4,
                            cur.cs4, cur.res, cur.have_res, cur.full);
        }
        __syncwarp(); // staging is rewritten by the next unit
      }
      if (c_first >= BN) { // a warp without columns (EW = 8, BN = 32) still releases the accumulators
        for (int k = 0; k < my_items; k++) {
          mbar_wait(&acc_full[k & 1], (uint32_t)(k >> 1) & 1);
          mbar_arrive(&acc_empty[k & 1]);
        }
      }
    } else {
    uint32_t li = 0;
    for (int i = blockIdx.x; i < items.total; i += gridDim.x, li++) {
      int m0, n0, kb_begin, nkb, z;
      int tile;
      items.decode(i, m0, n0, kb_begin, nkb, z, tile, BN);
      const int b = (int)(li & 1);
      // row state of the 8 rows this lane touches: bit j = in range, bit 8+j = live (not masked)
      uint32_t rows = 0;
      int mq = 0, mr = 0; // see the epilogue above
      if (e.rowmask) {
        const int mb = m0 + q * 32 + rsub;
        mq = mb / e.mask_div;
        mr = mb - mq * e.mask_div;
      }
#pragma unroll
      for (int j = 0; j < 8; j++) {
        const int m = m0 + q * 32 + j * 4 + rsub;
        if (m < M) {
          rows |= 1u << j;
          if (!e.rowmask || e.rowmask[mq])
            rows |= 0x100u << j;
        }
        if (e.rowmask) {
          mr += 4;
          while (mr >= e.mask_div) {
            mr -= e.mask_div;
            mq++;
          }
        }
      }
      const bool full = __all_sync(0xffffffffu, (rows & 0xFFu) == 0xFFu);
      float *tile_ws = g.ws + ((long)tile * S) * (long)(TC_BM * BN);
      float *mine = tile_ws + (long)z * TC_BM * BN;
      mbar_wait(&acc_full[b], (uint32_t)(li >> 1) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (et == 0 && li == 0)
        TC_STAMP(41);
      if (items.cluster)
        break; // one item per CTA; the accumulator is reduced across the cluster below
#pragma unroll 1
      for (int c0 = c_first; c0 < BN; c0 += C_STEP) {
        const int n = n0 + c0 + cf;
        const bool nok = n < N;
        uint32_t v[32];
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(b * BN + c0);
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
        // operands of this chunk that depend on the column only / the residual rows are
        // requested while the accumulator load is in flight
        float4 res[8];
        const bool have_res = (S == 1) && e.residual != nullptr;
        if (have_res) {
#pragma unroll
          for (int j = 0; j < 8; j++) {
            const int m = m0 + q * 32 + j * 4 + rsub;
            res[j] = (nok && ((rows >> j) & 1)) ? *reinterpret_cast<const float4 *>(e.residual + (long)m * e.ldr + n)
                                                : make_float4(0.f, 0.f, 0.f, 0.f);
          }
        }
        float4 bias4 = make_float4(0.f, 0.f, 0.f, 0.f), cs4 = make_float4(1.f, 1.f, 1.f, 1.f);
        if (S == 1 && nok) {
          if (e.bias)
            bias4 = __ldg(reinterpret_cast<const float4 *>(e.bias + n));
          if (e.colscale)
            cs4 = __ldg(reinterpret_cast<const float4 *>(e.colscale + n));
        }
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          sts128(stg_addr + (uint32_t)(lane * SM::STG_LD + j) * 4, v[j], v[j + 1], v[j + 2], v[j + 3]);
        __syncwarp();
        const uint32_t rd_addr = stg_addr + (uint32_t)(rsub * SM::STG_LD + cf) * 4;
        if (S > 1) {
#pragma unroll
          for (int j = 0; j < 8; j++) {
            const int row = j * 4 + rsub;
            *reinterpret_cast<float4 *>(mine + (long)(q * 32 + row) * BN + c0 + cf) =
                lds128(rd_addr + j * 4 * SM::STG_LD * 4);
          }
        } else if (nok) {
          const long m_first = m0 + q * 32 + rsub;
          epi_rows_dispatch(items.epi_kind, e, rd_addr, SM::STG_LD * 4, m_first, n, rows, bias4, cs4, res, have_res,
                            full);
        }
        __syncwarp(); // staging is rewritten by the next chunk
      }
      // this accumulator buffer may be overwritten by the MMA warp from here on
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      mbar_arrive(&acc_empty[b]);
      if (et == 0 && li == 0)
        TC_STAMP(42);
      if (ATMA && et == 0 && li < 32)
        TC_STAMP(48 + li); // per-item epilogue completion (the converter slots are free)
      if (S > 1) {
        // split-K: the last CTA to publish its partial tile sums all partials in split
        // order (deterministic) and runs the epilogue.  One thread publishes for the CTA:
        // the named barrier orders every epilogue thread's partial stores before its
        // gpu-scope fence + counter increment (cumulativity), and after them on the reader.
        asm volatile("bar.sync 1, %0;" ::"n"(EW * 32) : "memory");
        if (et == 0) {
          __threadfence();
          const int prev = atomicAdd(g.counters + tile, 1);
          const int last = (prev == S - 1);
          if (last)
            g.counters[tile] = 0;
          __threadfence();
          *s_last = last;
        }
        asm volatile("bar.sync 1, %0;" ::"n"(EW * 32) : "memory");
        const bool do_epilogue = *s_last != 0;
        asm volatile("bar.sync 1, %0;" ::"n"(EW * 32) : "memory"); // s_last is rewritten by the next item
        if (do_epilogue) {
#pragma unroll 1
          for (int c0 = c_first; c0 < BN; c0 += C_STEP) {
            const int n = n0 + c0 + cf;
#pragma unroll 1
            for (int j0 = 0; j0 < 8; j0 += 4) {
              float4 acc[4];
#pragma unroll
              for (int jj = 0; jj < 4; jj++)
                acc[jj] = make_float4(0.f, 0.f, 0.f, 0.f);
              // 16 partial loads in flight per lane (4 rows x 4 splits), summed in split order
              for (int sp0 = 0; sp0 < S; sp0 += 4) {
                float4 u[4][4];
#pragma unroll
                for (int jj = 0; jj < 4; jj++) {
                  const float *src = tile_ws + (long)(q * 32 + (j0 + jj) * 4 + rsub) * BN + c0 + cf;
#pragma unroll
                  for (int t = 0; t < 4; t++)
                    u[jj][t] = sp0 + t < S ? __ldcg(reinterpret_cast<const float4 *>(src + (long)(sp0 + t) * TC_BM * BN))
                                           : make_float4(0.f, 0.f, 0.f, 0.f);
                }
#pragma unroll
                for (int jj = 0; jj < 4; jj++)
#pragma unroll
                  for (int t = 0; t < 4; t++) {
                    acc[jj].x += u[jj][t].x; acc[jj].y += u[jj][t].y;
                    acc[jj].z += u[jj][t].z; acc[jj].w += u[jj][t].w;
                  }
              }
#pragma unroll
              for (int jj = 0; jj < 4; jj++) {
                const int j = j0 + jj;
                const int m = m0 + q * 32 + j * 4 + rsub;
                if (m < M && n < N) {
                  float4 o = epi_vec4(e, acc[jj], m, n);
                  if (!((rows >> (8 + j)) & 1))
                    o = make_float4(0.f, 0.f, 0.f, 0.f);
                  if (e.C)
                    *reinterpret_cast<float4 *>(e.C + (long)m * e.ldc + n) = o;
                  if (e.Cs.hi)
                    split_store4(e.Cs, m, n, o);
                  if (e.C2 || e.C2s.hi) {
                    const float4 u2 = elu4(o);
                    if (e.C2)
                      *reinterpret_cast<float4 *>(e.C2 + (long)m * e.ldc2 + n) = u2;
                    if (e.C2s.hi)
                      split_store4(e.C2s, m, n, u2);
                  }
                }
              }
            }
          }
        }
      }
    }
    } // S > 1 / cluster path
  }
  if (items.cluster) {
    // ===== split-K across a cluster, distributed reduction.  The S CTAs of a tile each own a
    // band of its 128 rows (rows_per = ceil(128 / S)).  Every CTA moves its partial accumulator
    // TMEM -> (per-warp transpose) -> the OWNER's shared memory through DSMEM (the operand ring
    // is idle by then), band by band; after the second cluster barrier each CTA sums the S
    // partials of its own band in split order (deterministic) and runs the fused epilogue on
    // it.  No partials through L2, no atomics, no serial leader: the exchange is one tile's
    // worth of bytes per CTA whatever S is, and the epilogue work is split S ways. =====
    __syncwarp();
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    int m0, n0, kb_begin, nkb, z, tile;
    items.decode((int)blockIdx.x, m0, n0, kb_begin, nkb, z, tile, BN);
    const bool epi_warp = warp >= 2 && warp < 2 + EW;
    const int q = warp & 3;
    const int c_first = EW == 8 ? ((warp - 2) >> 2) * 32 : 0;
    constexpr int C_STEP = EW == 8 ? 64 : 32;
    const int rows_per = (TC_BM + S - 1) / S;
    if (epi_warp) {
      const uint32_t my_base = smem_u32(smem);
      const uint32_t stg_addr_c = smem_u32(stg_base + (warp - 2) * SM::STG_WARP);
#pragma unroll 1
      for (int c0 = c_first; c0 < BN; c0 += C_STEP) {
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
        // transpose through the warp's staging buffer so that every remote store instruction
        // covers 4 rows x 128 contiguous bytes (a row-per-lane store is 32 separate 16-byte
        // DSMEM transactions and ran at ~4 GB/s)
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          sts128(stg_addr_c + (uint32_t)(lane * SM::STG_LD + j) * 4, v[j], v[j + 1], v[j + 2], v[j + 3]);
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 8; j++) {
          const int row = j * 4 + (lane >> 3);
          const int r = q * 32 + row;            // row of the tile
          const int owner = r / rows_per, rloc = r - owner * rows_per;
          const float4 a = lds128(stg_addr_c + (uint32_t)(row * SM::STG_LD + (lane & 7) * 4) * 4);
          uint32_t owner_base;
          asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(owner_base) : "r"(my_base), "r"(owner));
          const uint32_t dst = owner_base + (uint32_t)((((z * rows_per + rloc) * BN) + c0 + (lane & 7) * 4) * 4);
          asm volatile("st.shared::cluster.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(dst), "f"(a.x), "f"(a.y), "f"(a.z), "f"(a.w)
                       : "memory");
        }
        __syncwarp();
      }
    }
    __syncwarp();
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    if (epi_warp) {
      const GemmEpi &e = g.epi;
      const int r_begin = z * rows_per;
      const int rows_here = min(TC_BM, r_begin + rows_per) - r_begin;
      const uint32_t slots = smem_u32(smem);
      constexpr int C4 = BN / 4;
      const int total = rows_here * C4;
      for (int idx = threadIdx.x - 64; idx < total; idx += EW * 32) {
        const int rloc = idx / C4, c4 = idx - rloc * C4;
        const int m = m0 + r_begin + rloc, n = n0 + c4 * 4;
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int sp = 0; sp < S; sp++) {
          const float4 u = lds128(slots + (uint32_t)((((sp * rows_per + rloc) * BN) + c4 * 4) * 4));
          a.x += u.x; a.y += u.y; a.z += u.z; a.w += u.w;
        }
        if (m < M && n < N) {
          float4 o = epi_vec4(e, a, m, n);
          if (e.rowmask && !e.rowmask[m / e.mask_div])
            o = make_float4(0.f, 0.f, 0.f, 0.f);
          if (e.C)
            *reinterpret_cast<float4 *>(e.C + (long)m * e.ldc + n) = o;
          if (e.Cs.hi)
            split_store4(e.Cs, m, n, o);
          if (e.C2 || e.C2s.hi) {
            const float4 u2 = elu4(o);
            if (e.C2)
              *reinterpret_cast<float4 *>(e.C2 + (long)m * e.ldc2 + n) = u2;
            if (e.C2s.hi)
              split_store4(e.C2s, m, n, u2);
          }
        }
      }
    }
  }
  if (threadIdx.x == 64)
    TC_STAMP(43);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x == 0)
    TC_STAMP(44);
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

// How many clusters of `cs` CTAs of this kernel instance can be resident at once.  A GPC holds
// ~18 SMs and a cluster never spans GPCs, so "CTAs <= SMs" is not enough: 48 clusters of 3 do
// not fit although 144 <= 148, and the launch would run in two waves.
template <int BN, int STAGES, bool ATMA, int EW>
int tc_max_clusters(int cs) {
  using SM = TcSmem<BN, STAGES, EW>;
  static int cache[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  if (cs < 2 || cs > 8)
    return 0;
  if (cache[cs] == 0) {
    cudaFuncSetAttribute(tc_gemm_kernel<BN, STAGES, ATMA, EW>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM::BYTES);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(cs * 8));
    cfg.blockDim = dim3(tc_threads(EW, ATMA));
    cfg.dynamicSmemBytes = SM::BYTES;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)cs;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, tc_gemm_kernel<BN, STAGES, ATMA, EW>, &cfg) != cudaSuccess) {
      cudaGetLastError();
      n = -1;
    }
    cache[cs] = n > 0 ? n : -1;
  }
  return cache[cs] > 0 ? cache[cs] : 0;
}
// The same for the TMA-fed instance the tile width maps to (cost model + launch agree on it).
int tc_cluster_capacity(int bn, int cs) {
  switch (bn) {
  case 256: return tc_max_clusters<256, 2, true, 4>(cs);
  case 128: return tc_max_clusters<128, 3, true, 4>(cs);
  case 64: return tc_max_clusters<64, 4, true, 4>(cs);
  default: return tc_max_clusters<32, 4, true, 4>(cs);
  }
}

template <int BN, int STAGES, bool ATMA, int EW = 4, bool WLO = true>
int launch_tc_a(const TcWeight &w, const GemmArgs &g, int S, cudaStream_t st) {
  using SM = TcSmem<BN, STAGES, EW>;
  static volatile bool configured[64] = {};
  volatile bool &conf = device_flag(configured);
  if (!conf) {
    if (cudaFuncSetAttribute(tc_gemm_kernel<BN, STAGES, ATMA, EW, WLO>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM::BYTES) != cudaSuccess)
      return -1;
    conf = true;
  }
  const CUtensorMap *map = reinterpret_cast<const CUtensorMap *>(
      BN == 256 ? w.map256 : (BN == 128 ? w.map128 : (BN == 64 ? w.map64 : w.map32)));
  TcItems it;
  it.ntn = (g.N + BN - 1) / BN;
  const long tiles = (long)it.ntn * ((g.M + TC_BM - 1) / TC_BM);
  it.nkb_all = (g.K + TC_BK - 1) / TC_BK;
  if (S < 1 || !g.ws || !g.counters || tiles > g.n_counters)
    S = 1;
  if (S > it.nkb_all)
    S = it.nkb_all;
  while (S > 1 && tiles * S * (long)TC_BM * BN > g.ws_floats)
    S--;
  it.kb_per = (it.nkb_all + S - 1) / S;
  S = (it.nkb_all + it.kb_per - 1) / it.kb_per; // every split owns >= 1 k-block
  it.S = S;
  it.total = (int)(tiles * S);
  it.epi_kind = epi_kind_of(g.epi);
  // split-K through a cluster (DSMEM) when the S partial tiles fit the leader's operand ring
  static const bool no_cluster = getenv("PTTS_TC_NO_CLUSTER") != nullptr;
  it.cluster = (ATMA && !no_cluster && S > 1 && S <= 8 && it.total <= tc_num_sms() &&
                (long)S * ((TC_BM + S - 1) / S) * BN * 4 <= (long)STAGES * SM::STAGE &&
                tiles <= tc_max_clusters<BN, STAGES, ATMA, EW>(S))
                   ? 1
                   : 0;
  const int grid = (int)(it.total < tc_num_sms() ? it.total : tc_num_sms());
  const CUtensorMap *am = ATMA ? reinterpret_cast<const CUtensorMap *>(g.amap) : map;
  launch_kc(tc_gemm_kernel<BN, STAGES, ATMA, EW, WLO>, dim3(grid), dim3(tc_threads(EW, ATMA)), SM::BYTES, st,
            it.cluster ? S : 1, *map, *am, g, w.N, it);
  return 1;
}
template <int BN, int STAGES>
int launch_tc(const TcWeight &w, const GemmArgs &g, int S, cudaStream_t st) {
  if (g.amap && w.lo_zero)
    return launch_tc_a<BN, STAGES, true, 4, false>(w, g, S, st);
  return g.amap ? launch_tc_a<BN, STAGES, true>(w, g, S, st) : launch_tc_a<BN, STAGES, false>(w, g, S, st);
}
// Short-K TMA-fed shapes (the SEANet tail): the epilogue is the bottleneck, so it gets
// eight warps and the operand ring one stage less.
template <int BN, int STAGES>
int launch_tc_ew8(const TcWeight &w, const GemmArgs &g, int S, cudaStream_t st) {
  return launch_tc_a<BN, STAGES, true, 8>(w, g, S, st);
}

// Tile width and split-K for one GEMM.  Wide problems (many m tiles) take the widest
// tile that divides N (fewest A re-reads); skinny ones (the AR step, M <= 512) are
// bound by what one SM can pull through L2, so the choice minimises the bytes a CTA
// streams, (128 + BN) * 256 B per k-block, over one wave of CTAs.
int force_bn = -1, force_s = 0; // PTTS_TC_FORCE="bn,s" or tc_force_tile(): sweeps only

void choose_tile(const GemmArgs &g, int &bn, int &S) {
  if (force_bn < 0) {
    force_bn = 0;
    if (const char *f = getenv("PTTS_TC_FORCE"))
      sscanf(f, "%d,%d", &force_bn, &force_s);
  }
  const long mt = (g.M + TC_BM - 1) / TC_BM;
  const int nkb = (g.K + TC_BK - 1) / TC_BK;
  const int sms = tc_num_sms();
  const int cand[4] = {256, 128, 64, 32};
  if (force_bn > 0) {
    bn = force_bn;
    S = force_s > 0 ? force_s : 1;
    while (bn > 32 && bn / 2 >= g.N)
      bn /= 2;
    return;
  }
  // Cost model (us), fitted to the forced (BN, S) sweep of the AR-step shapes on B200
  // (tools/skinny_sweep.py, profiles/skinny_sweep_r02.txt):
  //  * ~6 us per launch are fixed (launch gap, barrier init + TMEM alloc, first operand round trip,
  //    teardown);
  //  * a CTA ingests its operands at ~95 GB/s (L2 -> SM), so a k-block costs (128 + BN) * 256 B /
  //    95 GB/s, but never less than the ring's latency floor: 0.58 us with 3-4 stages, 1.3 us with
  //    the two 96 KB stages of BN = 256 (and 1.3 us when A goes through the converter warps);
  //  * the whole grid cannot pull more than ~11 TB/s through L2;
  //  * the epilogue costs ~0.7 us per 32-column chunk; persistent CTAs overlap it with the next
  //    item's main loop;
  //  * split-K through a cluster (distributed reduction): two cluster barriers + one tile of DSMEM
  //    traffic per CTA (0.9 us per 64 columns) + 1/S of the epilogue; through L2 (more CTAs than
  //    SMs, or S > 8): 5 us + 2.5 us per 64 columns + 1 us per 64 columns and split.
  const bool atma = g.amap != nullptr;
  double best = 1e30;
  bn = 32;
  S = 1;
  for (int c = 0; c < 4; c++) {
    const int b = cand[c];
    if (b / 2 >= g.N && b > 32)
      continue; // more than half the tile would be padding
    const long ntn = (g.N + b - 1) / b;
    const long tiles = mt * ntn;
    const int smax = tiles >= sms ? 1 : 8;
    for (int s = 1; s <= smax && s <= nkb; s++) {
      const int kb_per = (nkb + s - 1) / s;
      if ((long)kb_per * (s - 1) >= nkb)
        continue;
      const long items = tiles * s;
      const long per_cta = (items + sms - 1) / sms; // items a CTA runs back to back
      const double kb_bytes = (128.0 + b) * 256.0;
      double t_kb = !atma ? 1.3 : (b == 256 ? 1.3 : 0.58);
      if (kb_bytes / 95e3 > t_kb)
        t_kb = kb_bytes / 95e3;
      const double t_mma = 128.0 * b * 64.0 * 6.0 / 9.0e6; // ~9 TFLOP/s issued per SM
      if (t_mma > t_kb)
        t_kb = t_mma;
      const double t_loop = t_kb * kb_per;
      const double t_epi = 0.7 * (b / 32.0);
      double t = 6.0 + per_cta * t_loop + (s == 1 ? t_epi : 0.0) +
                 (per_cta > 1 ? (per_cta - 1) * (t_epi > t_loop ? t_epi - t_loop : 0.0) : 0.0);
      const double l2_floor = (double)items * kb_per * kb_bytes / 11.0e6;
      if (6.0 + l2_floor > t)
        t = 6.0 + l2_floor;
      if (s > 1) {
        const long ring = b == 32 ? 4L * 40960 : 196608L;
        const bool cl = atma && s <= 8 && items <= sms && (long)s * ((128 + s - 1) / s) * b * 4 <= ring &&
                        tiles <= tc_cluster_capacity(b, s);
        t += cl ? 2.5 + (b / 64.0) * 0.9 + t_epi / s : 5.0 + (b / 64.0) * (2.5 + 1.0 * s) + t_epi;
      }
      if (t < best) {
        best = t;
        bn = b;
        S = s;
      }
    }
  }
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
  const int boxes[4] = {32, 64, 128, 256};
  void *maps[4] = {w->map32, w->map64, w->map128, w->map256};
  for (int i = 0; i < 4; i++) {
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

int tc_act_map_init(void *map, const void *base, long rows_total, int K, int cin, int taps) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc)
    return -1;
  CUresult r;
  cuuint32_t estr[3] = {1, 1, 1};
  if (cin > 0) {
    // {cin, taps, rows}: tap t of row m is row m + t of the plane (A rows overlap)
    cuuint64_t dims[3] = {(cuuint64_t)cin, (cuuint64_t)taps, (cuuint64_t)rows_total};
    cuuint64_t strides[2] = {(cuuint64_t)cin * 2, (cuuint64_t)cin * 2};
    cuuint32_t box[3] = {(cuuint32_t)TC_BK, 1, (cuuint32_t)TC_BM};
    r = enc(reinterpret_cast<CUtensorMap *>(map), CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void *>(base), dims,
            strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  } else {
    cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows_total};
    cuuint64_t strides[1] = {(cuuint64_t)K * 2};
    cuuint32_t box[2] = {(cuuint32_t)TC_BK, (cuuint32_t)TC_BM};
    r = enc(reinterpret_cast<CUtensorMap *>(map), CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void *>(base), dims,
            strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  }
  return r == CUDA_SUCCESS ? 0 : -2;
}

void tc_force_tile(int bn, int s) {
  force_bn = bn > 0 ? bn : 0;
  force_s = s;
}

bool tc_gemm_supported(const GemmArgs &g) {
  if (g.M < 1 || g.N < 32 || g.N % 16 != 0 || g.K < 32 || g.K % 8 != 0)
    return false;
  if (g.amap)
    return true;
  return g.A != nullptr && g.lda % 4 == 0 && (reinterpret_cast<uintptr_t>(g.A) & 15) == 0;
}

int tc_gemm(const TcWeight &w, const GemmArgs &g, cudaStream_t st, const char **variant) {
  int bn, S;
  choose_tile(g, bn, S);
  static const bool no_ew8 = getenv("PTTS_TC_NO_EW8") != nullptr;
  if (g.amap && g.K <= 512 && g.M >= 4096 && g.N >= 64 && g.N <= 256 && S == 1 && !no_ew8) {
    if (bn > 128)
      bn = 128; // two n tiles re-read A from L2, but the epilogue gets 8 warps
    if (bn == 128) {
      *variant = "tcgen05_128x128_ew8";
      return launch_tc_ew8<128, 2>(w, g, S, st);
    }
    *variant = "tcgen05_128x64_ew8";
    return launch_tc_ew8<64, 3>(w, g, S, st);
  }
  const bool w1 = g.amap && w.lo_zero; // bf16-exact weight: hi plane only
  switch (bn) {
  case 256:
    *variant = w1 ? "tcgen05_128x256_bf16w" : "tcgen05_128x256";
    return launch_tc<256, 2>(w, g, S, st);
  case 128:
    *variant = w1 ? "tcgen05_128x128_bf16w" : "tcgen05_128x128";
    return launch_tc<128, 3>(w, g, S, st);
  case 64:
    *variant = w1 ? "tcgen05_128x64_bf16w" : "tcgen05_128x64";
    return launch_tc<64, 4>(w, g, S, st);
  default:
    *variant = w1 ? "tcgen05_128x32_bf16w" : "tcgen05_128x32";
    return launch_tc<32, 4>(w, g, S, st);
  }
}

} // namespace ptts
