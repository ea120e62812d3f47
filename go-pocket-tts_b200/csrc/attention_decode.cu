// attention_decode.cu — decode attention over the paged fp16 KV cache with the K and V rows staged
// through shared memory by cp.async (attention.go:427-445 + rope.go:81-105 + the KV append of
// flow_lm.go:238-299 for one new token per sequence; same fusion as attention_paged_decode_kernel).
//
// Built on the hypothesis that attention_paged_decode_kernel (rows straight into registers, 2 KB in flight
// per pass of 16 keys) was latency-bound at 41-52 % of the HBM peak: here the CTA requests its whole K and V
// stream up front — 64-key tiles (8 KB) into a ring of four buffers, one cp.async group per tile, no
// registers held — and the arithmetic runs from shared memory behind it.  Thread-to-element assignment,
// summation order and the softmax are those of attend_row (attend.cuh), so the output is bit-identical to
// the register kernel's; the new token's K / V go to the cache first and are fetched back like every other row.
//
// Measured (profiles/README_r02.md): the same 69 us per launch at B = 512 — the hypothesis was wrong.  ncu
// shows the kernel ISSUE-bound (issue slots 67 % busy, DRAM 29 %): a third of its instructions were integer
// divisions by the run-time page size in the key -> (page, row) arithmetic.  With shifts this kernel reaches
// 64 us, and the same fix in the register kernel 57 us (4.3 TB/s at B = 1024), which also leaves the SM's
// shared memory to the next GEMM's CTAs.  So this variant is opt-in (PTTS_ATTN_STAGED=1); it stays as the
// parity-tested record of the experiment.
#include "attend.cuh"
#include "launch.cuh"

namespace ptts {

namespace {

constexpr int DT_KEYS = 64;                   // keys per tile
constexpr int DT_BYTES = DT_KEYS * HD * 2;    // 8 KB of fp16 rows
constexpr int DT_NB = 4;                      // ring buffers (a power of two)

__device__ __forceinline__ void cp_async16(void *dst, const void *src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src)
               : "memory");
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

__global__ void __launch_bounds__(ATT_THREADS)
    attention_paged_decode_staged_kernel(const float *qkv, long ldq, float *out, long ldo, const int32_t *row_slot,
                                         const int32_t *row_pos, const float *cosT, const float *sinT, KvGeom kv,
                                         int layer, int score_floats, int pt_shift, SplitOut sp) {
  PDL_TRIGGER();
  PDL_WAIT();
  extern __shared__ __align__(16) float smem[]; // scores[score_floats] | DT_NB tiles
  __shared__ float red[ATT_THREADS / 32];
  __shared__ float4 pv[8 * 16];
  __shared__ __align__(16) float qs[HD];
  __shared__ int32_t s_pt[KV_MAX_SEQ_PAGES];
  const int r = blockIdx.x, h = blockIdx.y, tid = threadIdx.x;
  const int slot = row_slot[r], pos = row_pos[r];
  float *o = out ? out + (long)r * ldo + h * HD : nullptr;
  if (pos < 0) {
    if (tid < 16) {
      const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
      if (o)
        *reinterpret_cast<float4 *>(o + tid * 4) = z;
      if (sp.hi)
        split_store4(sp, r, h * HD + tid * 4, z);
    }
    return;
  }
  const int nkeys = pos + 1;
  const int32_t *pt = kv.pagetab + (long)slot * kv.max_pages;
  const int PT = kv.page_tokens, D = kv.n_heads * HD;
  const long page_stride = (long)kv.n_layers * 2 * kv.n_heads * PT * HD;
  const long koff = ((long)(layer * 2 + 0) * kv.n_heads + h) * PT * HD;
  const long voff = ((long)(layer * 2 + 1) * kv.n_heads + h) * PT * HD;
  __half *pool = static_cast<__half *>(kv.pool);
  float *scores = smem;
  unsigned char *tiles = reinterpret_cast<unsigned char *>(smem + score_floats);
  for (int i = tid; (i << pt_shift) < nkeys; i += ATT_THREADS)
    s_pt[i] = pt[i];
  {
    const float *row = qkv + (long)r * ldq;
    __half *page = pool + (long)pt[pos >> pt_shift] * page_stride + (long)(pos & (PT - 1)) * HD;
    if (tid < HD) { // pairs (2j, 2j+1): threads 0..31 rotate q, 32..63 rotate k
      const int isk = tid >= HD / 2, j = tid & (HD / 2 - 1);
      const float *src = row + (isk ? D : 0) + h * HD + 2 * j;
      const float a = src[0], b = src[1];
      const float cj = cosT[(long)pos * (HD / 2) + j], sj = sinT[(long)pos * (HD / 2) + j];
      const float o0 = rope_even(a, b, cj, sj);
      const float o1 = rope_odd(a, b, cj, sj);
      if (isk) {
        kv_store(page + koff + 2 * j, o0);
        kv_store(page + koff + 2 * j + 1, o1);
      } else {
        qs[2 * j] = o0;
        qs[2 * j + 1] = o1;
      }
    } else { // threads 64..127: the head's v
      const int d = tid - HD;
      kv_store(page + voff + d, row[2 * D + h * HD + d]);
    }
  }
  __syncthreads(); // q and the page ids in shared memory; this CTA's own k/v stores precede the copies below
  // tile i of the stream: K tiles 0..T-1, then V tiles 0..T-1; tile i lives in buffer i % DT_NB
  const int T = (nkeys + DT_KEYS - 1) / DT_KEYS, ntiles = 2 * T;
  // 512 x 16 B per tile, 4 per thread: piece u of a thread is key (tid >> 3) + 16 u of the tile, bytes
  // [16 (tid & 7), +16) of its row.  The page size is a power of two (shifts: the address arithmetic is what
  // this kernel issues most).  Always one commit group per call (possibly empty).
  const uint32_t tiles_s = (uint32_t)__cvta_generic_to_shared(tiles) + (uint32_t)tid * 16;
  const __half *src_k = pool + koff + (tid & 7) * 8, *src_v = pool + voff + (tid & 7) * 8;
  const int pmask = PT - 1, jrow = tid >> 3;
  auto issue = [&](int i) {
    if (i < ntiles) {
      const int t = i < T ? i : i - T;
      const __half *src = i < T ? src_k : src_v;
      const uint32_t dst = tiles_s + (uint32_t)(i & (DT_NB - 1)) * DT_BYTES;
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const int j = t * DT_KEYS + jrow + 16 * u;
        if (j < nkeys) {
          const __half *g = src + (long)s_pt[j >> pt_shift] * page_stride + ((j & pmask) << 6);
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + (uint32_t)u * 2048), "l"(g) : "memory");
        }
      }
    }
    cp_commit();
  };
#pragma unroll
  for (int i = 0; i < DT_NB; i++)
    issue(i);
  const float scale = 0.125f; // 1/sqrt(64)
  // ---- scores (attend_row phase 1: lane `part` of a key's 8 lanes owns elements [8 part, 8 part + 8)) ----
  const int part = tid & 7, kslot = tid >> 3;
  const float4 q0 = *reinterpret_cast<const float4 *>(qs + part * 8);
  const float4 q1 = *reinterpret_cast<const float4 *>(qs + part * 8 + 4);
  float lmax = -INFINITY;
  for (int t = 0; t < T; t++) {
    cp_wait<DT_NB - 1>();
    __syncthreads();
    const __half *buf = reinterpret_cast<const __half *>(tiles + (size_t)(t & (DT_NB - 1)) * DT_BYTES);
#pragma unroll
    for (int p = 0; p < DT_KEYS / 16; p++) {
      const int jj = p * 16 + kslot, j = t * DT_KEYS + jj;
      float s = 0.0f;
      if (j < nkeys) {
        float k[8];
        kv_load8(buf + jj * HD + part * 8, k);
        s = q0.x * k[0];
        s = fmaf(q0.y, k[1], s);
        s = fmaf(q0.z, k[2], s);
        s = fmaf(q0.w, k[3], s);
        s = fmaf(q1.x, k[4], s);
        s = fmaf(q1.y, k[5], s);
        s = fmaf(q1.z, k[6], s);
        s = fmaf(q1.w, k[7], s);
      }
      s += __shfl_xor_sync(0xffffffffu, s, 1);
      s += __shfl_xor_sync(0xffffffffu, s, 2);
      s += __shfl_xor_sync(0xffffffffu, s, 4);
      if (j < nkeys) {
        s *= scale;
        if (part == 0)
          scores[j] = s;
        lmax = fmaxf(lmax, s);
      }
    }
    if (t + DT_NB < ntiles)
      __syncthreads(); // everybody is done with this buffer before tile t + DT_NB overwrites it
    issue(t + DT_NB);
  }
  const float gmax = block_reduce_max(lmax, red); // includes __syncthreads: every score is written
  // ---- exp / sum (phase 2) ----
  float lsum = 0.0f;
  for (int j = tid; j < nkeys; j += ATT_THREADS) {
    const float e = expf(scores[j] - gmax);
    scores[j] = e;
    lsum += e;
  }
  const float gsum = block_reduce_sum(lsum, red);
  const float inv = 1.0f / gsum;
  // ---- P.V (phase 3: thread (grp, c4) adds keys grp, grp + 8, ... in ascending order) ----
  const int c4 = tid & 15, grp = tid >> 4;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int t = 0; t < T; t++) {
    cp_wait<DT_NB - 1>();
    __syncthreads();
    const __half *buf = reinterpret_cast<const __half *>(tiles + (size_t)((T + t) & (DT_NB - 1)) * DT_BYTES);
#pragma unroll
    for (int i = 0; i < DT_KEYS / 8; i++) {
      const int jj = grp + 8 * i, j = t * DT_KEYS + jj;
      if (j < nkeys) {
        const float w = scores[j] * inv;
        const float4 v = kv_load4(buf + jj * HD + c4 * 4);
        acc.x = fmaf(w, v.x, acc.x);
        acc.y = fmaf(w, v.y, acc.y);
        acc.z = fmaf(w, v.z, acc.z);
        acc.w = fmaf(w, v.w, acc.w);
      }
    }
    if (T + t + DT_NB < ntiles)
      __syncthreads();
    issue(T + t + DT_NB);
  }
  pv[grp * 16 + c4] = acc;
  __syncthreads();
  if (tid < 16) {
    float4 rr = pv[tid];
#pragma unroll
    for (int g2 = 1; g2 < 8; g2++) {
      const float4 t4 = pv[g2 * 16 + tid];
      rr.x += t4.x; rr.y += t4.y; rr.z += t4.z; rr.w += t4.w;
    }
    if (o)
      *reinterpret_cast<float4 *>(o + tid * 4) = rr;
    if (sp.hi)
      split_store4(sp, r, h * HD + tid * 4, rr);
  }
}

} // namespace

// Returns 0 when the staged kernel does not apply (fp32 cache, page table longer than the staging array).
int attention_paged_decode_staged(const float *qkv, long ldq, float *out, long ldo, int rows, const int32_t *row_slot,
                                  const int32_t *row_pos, const float *cosT, const float *sinT, const KvGeom &kv,
                                  int layer, int max_keys, cudaStream_t st, SplitOut sp) {
  const int PT = kv.page_tokens;
  if (rows <= 0 || !kv.half || PT <= 0 || (PT & (PT - 1)) != 0 || (max_keys + PT - 1) / PT > KV_MAX_SEQ_PAGES)
    return 0;
  int pt_shift = 0;
  while ((1 << pt_shift) < PT)
    pt_shift++;
  const int score_floats = (max_keys + 15) / 16 * 16;
  const size_t sm = sizeof(float) * (size_t)score_floats + (size_t)DT_NB * DT_BYTES;
  if (sm > 48 * 1024)
    return 0;
  dim3 grid(rows, kv.n_heads);
  launch_k(attention_paged_decode_staged_kernel, grid, dim3(ATT_THREADS), sm, st, qkv, ldq, out, ldo, row_slot, row_pos,
           cosT, sinT, kv, layer, score_floats, pt_shift, sp);
  return 1;
}

} // namespace ptts
