// flow_head.cu — the batched flow-matching / LSD head of one AR frame as ONE persistent
// cooperative kernel (north star item 2; reference internal/native/flow_lm.go:262-353 over
// flow_net.go:314-356, :116-172, :205-239).
//
// For B live sequences and one LSD step the head is a dependent chain of row-local ops:
//   last = out_norm(x)                      eos = out_eos(last)          cur = noise
//   sy   = silu(cond_embed(last) + tvec)    xf  = input_proj(cur)
//   ada  = adaLN_all(sy)                    [B, 6*3C + 2C]
//   6 x { h = LN(xf)(1+scale)+shift ; h2 = silu(mlp0(h)) ; xf += gate * mlp2(h2) }
//   h = LN(xf)(1+fscale)+fshift ; cur += final(h)
// As separate launches that is 27 kernels per frame, most of them far below a microsecond of
// work.  Here every CTA (one per SM, all co-resident) walks the same list of phases; a phase is
// either a row-wise op (one warp per row) or a tcgen05 GEMM whose 128 x BN output tiles are
// dealt round-robin to the CTAs; phases are separated by a grid barrier (16 striped counters,
// bounded spins).  The tensor-core machinery is the one of tcgemm.cu: operands by TMA (128-byte
// swizzle) into a 3-stage ring that runs straight through the phase boundaries, split-bf16
// operands (3 products per k-step, fp32 TMEM accumulator, double-buffered), the epilogue warps
// read TMEM 32 columns at a time, transpose through shared memory and run the same fused
// epilogue (bias, SiLU, adaLN gate, residual, Euler update, split-plane outputs).  The weight
// tiles of the first item of the NEXT phase are requested before the CTA arrives at the
// barrier, so the weight stream does not stop at the phase boundaries.
#include "kernels.cuh"
#include "gemm_epi.cuh"
#include "split.cuh"
#include "tc_common.cuh"
#include "flow_head.cuh"
#include <cuda.h>
#include <stdio.h>

namespace ptts {

namespace {

constexpr int FH_THREADS = 192;      // warp 0 TMA, warp 1 MMA, warps 2..5 epilogue; all six run the row-wise phases
constexpr int FH_STAGES = 3;
constexpr int FH_STAGE = 65536;      // bytes per ring stage: A hi/lo 32 KB + W hi/lo up to 32 KB (BN = 128)
constexpr int FH_A_TILE = TC_BM * TC_BK * 2;
constexpr int FH_STG_LD = 36;
constexpr int FH_STG_WARP = 32 * FH_STG_LD * 4;
constexpr int FH_SMEM = FH_STAGES * FH_STAGE + 4 * FH_STG_WARP + 1024 + 256;
constexpr int FH_TMEM_COLS = 256;    // two accumulators of up to 128 columns

constexpr int GB_NCTR = 16, GB_STRIDE = 32;

struct Bar {
  unsigned *ctr;
  int *err;
  unsigned phase, mine;
  bool dead;
};
__device__ __forceinline__ void bar_arrive(Bar &b) {
  b.phase++;
  // generic-proxy stores of this CTA (epilogue outputs) are read by other CTAs' TMA loads
  asm volatile("fence.proxy.async;" ::: "memory");
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    asm volatile("red.relaxed.gpu.global.add.u32 [%0], 1;" ::"l"(b.ctr + GB_STRIDE * (blockIdx.x & (GB_NCTR - 1)))
                 : "memory");
  }
}
__device__ __forceinline__ void bar_wait(Bar &b) {
  if (threadIdx.x < GB_NCTR && !b.dead) {
    const unsigned target = b.phase * b.mine;
    const unsigned *p = b.ctr + GB_STRIDE * threadIdx.x;
    long spins = 0;
    for (;;) {
      unsigned v;
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
      if (v >= target)
        break;
      if (++spins > 4000000L || ((spins & 255) == 0 && *reinterpret_cast<volatile int *>(b.err) != 0)) {
        atomicExch(b.err, 1); // somebody is missing: give up instead of hanging the GPU
        break;
      }
    }
  }
  __syncthreads();
  asm volatile("fence.proxy.async;" ::: "memory");
  if (!b.dead && *reinterpret_cast<volatile int *>(b.err) != 0)
    b.dead = true;
}

__device__ __forceinline__ float warp_sum(float s) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
    s += __shfl_xor_sync(0xffffffffu, s, o);
  return s;
}

// One row of NV*128 floats in registers: two-pass fp32 LayerNorm (linear.go:265-329), optional
// affine, optional adaLN modulate (tensor_util.go:175-193), fp32 and / or split-plane output.
template <int NV>
__device__ __forceinline__ void ln_row(const float *xr, int lane, const float *w, const float *b, float eps,
                                       const float *scale, const float *shift, float *out, const SplitOut &sp, long row,
                                       float4 (&o)[NV]) {
  constexpr int D = NV * 128;
  float4 v[NV];
#pragma unroll
  for (int j = 0; j < NV; j++)
    v[j] = __ldcg(reinterpret_cast<const float4 *>(xr + lane * 4 + 128 * j));
  float s = 0.0f;
#pragma unroll
  for (int j = 0; j < NV; j++)
    s += (v[j].x + v[j].y) + (v[j].z + v[j].w);
  const float mean = warp_sum(s) / (float)D;
  float q = 0.0f;
#pragma unroll
  for (int j = 0; j < NV; j++) {
    const float a = v[j].x - mean, c = v[j].y - mean, d = v[j].z - mean, e = v[j].w - mean;
    q += (a * a + c * c) + (d * d + e * e);
  }
  const float inv = 1.0f / sqrtf(warp_sum(q) / (float)D + eps);
#pragma unroll
  for (int j = 0; j < NV; j++) {
    const int i = lane * 4 + 128 * j;
    float n[4] = {(v[j].x - mean) * inv, (v[j].y - mean) * inv, (v[j].z - mean) * inv, (v[j].w - mean) * inv};
    if (w) {
      const float4 wv = __ldg(reinterpret_cast<const float4 *>(w + i));
      n[0] *= wv.x; n[1] *= wv.y; n[2] *= wv.z; n[3] *= wv.w;
    }
    if (b) {
      const float4 bv = __ldg(reinterpret_cast<const float4 *>(b + i));
      n[0] += bv.x; n[1] += bv.y; n[2] += bv.z; n[3] += bv.w;
    }
    if (scale) {
      const float4 sc = __ldcg(reinterpret_cast<const float4 *>(scale + i));
      const float4 sh = __ldcg(reinterpret_cast<const float4 *>(shift + i));
      n[0] = n[0] * (sc.x + 1.0f) + sh.x;
      n[1] = n[1] * (sc.y + 1.0f) + sh.y;
      n[2] = n[2] * (sc.z + 1.0f) + sh.z;
      n[3] = n[3] * (sc.w + 1.0f) + sh.w;
    }
    o[j] = make_float4(n[0], n[1], n[2], n[3]);
    if (out)
      *reinterpret_cast<float4 *>(out + i) = o[j];
    if (sp.hi)
      split_store4(sp, row, i, o[j]);
  }
}

// State of the tensor-core pipeline that lives across phases.
struct Pipe {
  uint8_t *smem;
  uint64_t *full, *empty, *acc_full, *acc_empty;
  uint32_t tmem_base;
  uint32_t it;  // k-blocks issued / consumed so far (ring position), per role
  uint32_t li;  // items so far (accumulator buffer), per role
};

// The W tiles of one k-block (hi rows at n0, lo rows at n_rows_hi + n0).
__device__ __forceinline__ void load_w(const FhGemm &g, uint8_t *st, uint64_t *bar, int k0, int n0) {
  const int wt = g.bn * TC_BK * 2;
  tma_load_2d(st + 2 * FH_A_TILE, g.wmap, bar, k0, n0);
  tma_load_2d(st + 2 * FH_A_TILE + wt, g.wmap, bar, k0, g.n_rows_hi + n0);
}
__device__ __forceinline__ void load_a(const FhGemm &g, uint8_t *st, uint64_t *bar, int k0, int m0) {
  tma_load_2d(st, g.amap, bar, k0, m0);
  tma_load_2d(st + FH_A_TILE, g.amap, bar, k0, g.a_lo_row + m0);
}

template <int BN>
__device__ __forceinline__ void mma_item(Pipe &p, int nkb) {
  constexpr uint32_t idesc = make_idesc(BN);
  const int b = (int)(p.li & 1);
  if (p.li >= 2) {
    mbar_wait(&p.acc_empty[b], (uint32_t)((p.li >> 1) - 1) & 1);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  }
  const uint32_t tmem_d = p.tmem_base + (uint32_t)(b * 128);
  for (int kb = 0; kb < nkb; kb++, p.it++) {
    const int s = (int)(p.it % FH_STAGES);
    mbar_wait(&p.full[s], (uint32_t)(p.it / FH_STAGES) & 1);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t a_hi = smem_u32(p.smem + s * FH_STAGE);
    const uint32_t a_lo = a_hi + FH_A_TILE;
    const uint32_t w_hi = a_hi + 2 * FH_A_TILE;
    const uint32_t w_lo = w_hi + BN * TC_BK * 2;
#pragma unroll
    for (int k = 0; k < TC_BK / 16; k++) {
      const uint32_t ko = k * 32;
      const uint64_t dah = make_smem_desc(a_hi + ko), dal = make_smem_desc(a_lo + ko);
      const uint64_t dwh = make_smem_desc(w_hi + ko), dwl = make_smem_desc(w_lo + ko);
      umma_bf16(tmem_d, dal, dwh, idesc, (kb | k) != 0);
      umma_bf16(tmem_d, dah, dwl, idesc, 1);
      umma_bf16(tmem_d, dah, dwh, idesc, 1);
    }
    umma_commit(&p.empty[s]);
  }
  umma_commit(&p.acc_full[b]);
  p.li++;
}

// Epilogue of one item by the four epilogue warps (warp q owns TMEM lanes 32 q .. 32 q + 31).
__device__ __forceinline__ void epi_item(Pipe &p, const FhGemm &g, int M, int m0, int n0, int warp, int lane,
                                         uint32_t stg_addr) {
  const GemmEpi &e = g.epi;
  const int q = warp & 3;
  const int b = (int)(p.li & 1);
  const int rsub = lane >> 3, cf = (lane & 7) * 4;
  uint32_t rows = 0;
#pragma unroll
  for (int j = 0; j < 8; j++) {
    const int m = m0 + q * 32 + j * 4 + rsub;
    if (m < M)
      rows |= (1u << j) | (0x100u << j);
  }
  const bool full = __all_sync(0xffffffffu, (rows & 0xFFu) == 0xFFu);
  mbar_wait(&p.acc_full[b], (uint32_t)(p.li >> 1) & 1);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
  for (int c0 = 0; c0 < g.bn; c0 += 32) {
    const int n = n0 + c0 + cf;
    const bool nok = n < g.N;
    uint32_t v[32];
    const uint32_t taddr = p.tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(b * 128 + c0);
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
    const bool have_res = e.residual != nullptr;
    if (have_res) {
#pragma unroll
      for (int j = 0; j < 8; j++) {
        const int m = m0 + q * 32 + j * 4 + rsub;
        res[j] = (nok && ((rows >> j) & 1)) ? __ldcg(reinterpret_cast<const float4 *>(e.residual + (long)m * e.ldr + n))
                                            : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
    float4 bias4 = make_float4(0.f, 0.f, 0.f, 0.f), cs4 = make_float4(1.f, 1.f, 1.f, 1.f);
    if (nok && e.bias)
      bias4 = __ldg(reinterpret_cast<const float4 *>(e.bias + n));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int j = 0; j < 32; j += 4)
      sts128(stg_addr + (uint32_t)(lane * FH_STG_LD + j) * 4, v[j], v[j + 1], v[j + 2], v[j + 3]);
    __syncwarp();
    if (nok) {
      const uint32_t rd_addr = stg_addr + (uint32_t)(rsub * FH_STG_LD + cf) * 4;
      const long m_first = m0 + q * 32 + rsub;
      epi_rows_dispatch(g.epi_kind, e, rd_addr, FH_STG_LD * 4, m_first, n, rows, bias4, cs4, res, have_res, full);
    }
    __syncwarp();
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  mbar_arrive(&p.acc_empty[b]);
  p.li++;
}

__global__ void __launch_bounds__(FH_THREADS, 1) flow_head_kernel(const FhParams *__restrict__ Pp, int B) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t *stg_base = smem + FH_STAGES * FH_STAGE;
  uint64_t *full = reinterpret_cast<uint64_t *>(stg_base + 4 * FH_STG_WARP);
  uint64_t *empty = full + FH_STAGES;
  uint64_t *acc_full = empty + FH_STAGES;
  uint64_t *acc_empty = acc_full + 2;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(acc_empty + 2);
  const FhParams &P = *Pp;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < FH_STAGES; s++) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int b = 0; b < 2; b++) {
      mbar_init(&acc_full[b], 1);
      mbar_init(&acc_empty[b], 4 * 32);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "n"(FH_TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  Pipe pipe;
  pipe.smem = smem;
  pipe.full = full;
  pipe.empty = empty;
  pipe.acc_full = acc_full;
  pipe.acc_empty = acc_empty;
  pipe.tmem_base = *tmem_slot;
  pipe.it = 0;
  pipe.li = 0;

  Bar bar;
  bar.ctr = P.barrier;
  bar.err = P.error;
  bar.phase = 0;
  bar.dead = false;
  bar.mine = (gridDim.x + GB_NCTR - 1 - (threadIdx.x & (GB_NCTR - 1))) / GB_NCTR;

  const int gwarp = blockIdx.x * (FH_THREADS / 32) + warp, nwarps = gridDim.x * (FH_THREADS / 32);
  const int mt = (B + TC_BM - 1) / TC_BM;
  const uint32_t stg_addr = smem_u32(stg_base + ((warp - 2) & 3) * FH_STG_WARP);

  // The TMA warp remembers which stages already hold the weight tiles of the next phase's first
  // items (requested before the barrier); their A tiles follow once the barrier has been passed.
  int pre_n = 0;  // k-blocks of the next GEMM whose W tiles are already in flight

  for (int ph = 0; ph < P.n_phases; ph++) {
    const FhPhase &F = P.phase[ph];
    if (F.kind != FH_GEMM && (P.debug & 4)) {
    } else if (F.kind == FH_ROWS_HEAD) {
      // ---- out_norm + EOS logit + cur = noise + xf = input_proj(cur) (flow_lm.go:262-283, flow_net.go:325) ----
      for (int r = gwarp; r < B; r += nwarps) {
        float4 o[8];
        ln_row<8>(P.x + (long)r * P.D, lane, P.on_w, P.on_b, P.on_eps, nullptr, nullptr, P.last + (long)r * P.D, P.s_last, r, o);
        float s = 0.0f;
#pragma unroll
        for (int j = 0; j < 8; j++) {
          const float4 w = __ldg(reinterpret_cast<const float4 *>(P.eos_w + lane * 4 + 128 * j));
          s = fmaf(o[j].x, w.x, s);
          s = fmaf(o[j].y, w.y, s);
          s = fmaf(o[j].z, w.z, s);
          s = fmaf(o[j].w, w.w, s);
        }
        s = warp_sum(s);
        if (lane == 0)
          P.eos[r] = s + P.eos_b;
        // cur = noise (ldim == 32: one element per lane), fp32 + split planes
        const float nz = P.noise[(long)r * P.ldim + lane];
        P.cur[(long)r * P.ldim + lane] = nz;
        split_store1(P.s_cur, r, lane, nz);
        // xf = input_proj(cur) + b: C outputs, K = 32; lane l owns outputs l, l + 32, ...
        for (int c = lane; c < P.C; c += 32) {
          const float *w = P.ip_w + (long)c * P.ldim;
          float a = 0.0f;
#pragma unroll 8
          for (int k = 0; k < 32; k++)
            a = fmaf(__shfl_sync(0xffffffffu, nz, k), __ldg(w + k), a);
          P.xf[(long)r * P.C + c] = a + (P.ip_b ? __ldg(P.ip_b + c) : 0.0f);
        }
      }
    } else if (F.kind == FH_ROWS_LN) {
      // ---- h = LN(xf)(1 + scale) + shift -> split planes (flow_net.go:116-172, :205-239) ----
      for (int r = gwarp; r < B; r += nwarps) {
        float4 o[4];
        ln_row<4>(P.xf + (long)r * P.C, lane, F.ln_w, F.ln_b, F.ln_eps, P.ada + (long)r * P.ada_n + F.scale_off,
                  P.ada + (long)r * P.ada_n + F.shift_off, nullptr, P.s_h, r, o);
      }
    } else {
      // ---- tcgen05 GEMM phase ----
      const FhGemm &g = F.g;
      const int ntn = (g.N + g.bn - 1) / g.bn;
      const int total = (P.debug & 2) ? 0 : mt * ntn;
      const int nkb = (g.K + TC_BK - 1) / TC_BK;
      if (warp == 0) {
        if (lane == 0) {
          for (int i = blockIdx.x; i < total; i += gridDim.x) {
            const int m0 = (i / ntn) * TC_BM, n0 = (i % ntn) * g.bn;
            const bool first = i == (int)blockIdx.x;
            for (int kb = 0; kb < nkb; kb++) {
              const uint32_t it = pipe.it + (uint32_t)kb;
              const int s = (int)(it % FH_STAGES);
              uint8_t *st = smem + s * FH_STAGE;
              const int k0 = kb * TC_BK;
              if (first && kb < pre_n) { // W already in flight, expect_tx already posted for the whole stage
                load_a(g, st, &full[s], k0, m0);
                continue;
              }
              if (it >= FH_STAGES)
                mbar_wait(&empty[s], (uint32_t)((it / FH_STAGES) - 1) & 1);
              mbar_arrive_expect_tx(&full[s], 2 * FH_A_TILE + 2 * g.bn * TC_BK * 2);
              load_a(g, st, &full[s], k0, m0);
              load_w(g, st, &full[s], k0, n0);
            }
            pipe.it += (uint32_t)nkb;
          }
          pre_n = 0;
          // weight tiles of this CTA's first item of the next GEMM phase, before the barrier
          int nx = ph + 1;
          while (nx < P.n_phases && P.phase[nx].kind != FH_GEMM)
            nx++;
          if (nx < P.n_phases) {
            const FhGemm &h = P.phase[nx].g;
            const int hntn = (h.N + h.bn - 1) / h.bn, htotal = (P.debug & 2) ? 0 : mt * hntn, hnkb = (h.K + TC_BK - 1) / TC_BK;
            if ((int)blockIdx.x < htotal) {
              const int n0 = ((int)blockIdx.x % hntn) * h.bn;
              const int can = hnkb < FH_STAGES ? hnkb : FH_STAGES;
              for (int kb = 0; kb < can; kb++) {
                const uint32_t it = pipe.it + (uint32_t)kb;
                const int s = (int)(it % FH_STAGES);
                if (it >= FH_STAGES) {
                  // only stages this phase has already drained can be reused without blocking on
                  // MMAs still to come: all of this phase's loads are issued, so waiting is safe
                  mbar_wait(&empty[s], (uint32_t)((it / FH_STAGES) - 1) & 1);
                }
                mbar_arrive_expect_tx(&full[s], 2 * FH_A_TILE + 2 * h.bn * TC_BK * 2);
                load_w(h, smem + s * FH_STAGE, &full[s], kb * TC_BK, n0);
                pre_n++;
              }
            }
          }
        }
      } else if (warp == 1) {
        if (lane == 0) {
          for (int i = blockIdx.x; i < total; i += gridDim.x) {
            if (g.bn == 128)
              mma_item<128>(pipe, nkb);
            else if (g.bn == 64)
              mma_item<64>(pipe, nkb);
            else
              mma_item<32>(pipe, nkb);
          }
        }
      } else {
        for (int i = blockIdx.x; i < total; i += gridDim.x) {
          const int m0 = (i / ntn) * TC_BM, n0 = (i % ntn) * g.bn;
          epi_item(pipe, g, B, m0, n0, warp, lane, stg_addr);
        }
      }
    }
    if (ph + 1 < P.n_phases && !F.no_barrier_after && !(P.debug & 1)) {
      bar_arrive(bar);
      bar_wait(bar);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(pipe.tmem_base), "n"(FH_TMEM_COLS) : "memory");
}

} // namespace

size_t flow_head_smem() { return FH_SMEM; }

int flow_head_grid() {
  int dev = 0, sms = 0, occ = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (cudaFuncSetAttribute(flow_head_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FH_SMEM) != cudaSuccess)
    return 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, flow_head_kernel, FH_THREADS, FH_SMEM) != cudaSuccess || occ < 1)
    return 0;
  return sms; // one CTA per SM
}

int flow_head_epi_kind(const GemmEpi &e) { return epi_kind_of(e); }

int flow_head_launch(const FhParams *dp, int grid, int B, unsigned *barrier, cudaStream_t st) {
  if (cudaMemsetAsync(barrier, 0, FH_BARRIER_BYTES, st) != cudaSuccess)
    return -1;
  void *args[] = {(void *)&dp, (void *)&B};
  if (cudaLaunchCooperativeKernel((void *)flow_head_kernel, dim3(grid), dim3(FH_THREADS), args, FH_SMEM, st) != cudaSuccess)
    return -1;
  return 2;
}

} // namespace ptts
