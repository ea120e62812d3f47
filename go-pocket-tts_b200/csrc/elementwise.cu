// elementwise.cu — row-wise and gather/scatter kernels of the hot path.
#include "kernels.cuh"
#include "split.cuh"
#include "launch.cuh"
#include <math.h>

namespace ptts {

static thread_local bool g_pdl = false;
bool pdl_enabled() { return g_pdl; }
void pdl_set(bool on) { g_pdl = on; }

// LayerNorm (+ optional adaLN modulate).  One warp per row, two-pass fp32
// statistics about the mean (reference: fp64 mean / biased variance,
// internal/native/linear.go:265-329; fp32 two-pass differs by ~1e-7 relative).
__global__ void __launch_bounds__(256)
    layernorm_kernel(const float *x, long ldx, float *out, long ldo, int rows,
                     int D, const float *w, const float *b, float eps,
                     const float *scale, const float *shift, long ldmod, SplitOut sp) {
  PDL_TRIGGER();
  PDL_WAIT();
  const int lane = threadIdx.x & 31;
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (row >= rows)
    return;
  const float *xr = x + (long)row * ldx;
  float s = 0.0f;
  for (int i = lane * 4; i < D; i += 128) {
    float4 v = *reinterpret_cast<const float4 *>(xr + i);
    s += (v.x + v.y) + (v.z + v.w);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
    s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mean = s / (float)D;
  float q = 0.0f;
  for (int i = lane * 4; i < D; i += 128) {
    float4 v = *reinterpret_cast<const float4 *>(xr + i);
    float a = v.x - mean, c = v.y - mean, d = v.z - mean, e = v.w - mean;
    q += (a * a + c * c) + (d * d + e * e);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
    q += __shfl_xor_sync(0xffffffffu, q, o);
  const float inv = 1.0f / sqrtf(q / (float)D + eps);
  float *orow = out + (long)row * ldo;
  for (int i = lane * 4; i < D; i += 128) {
    float4 v = *reinterpret_cast<const float4 *>(xr + i);
    float n[4] = {(v.x - mean) * inv, (v.y - mean) * inv, (v.z - mean) * inv,
                  (v.w - mean) * inv};
#pragma unroll
    for (int j = 0; j < 4; j++) {
      float t = n[j];
      if (w)
        t *= __ldg(w + i + j);
      if (b)
        t += __ldg(b + i + j);
      if (scale)
        t = t * (scale[(long)row * ldmod + i + j] + 1.0f) + shift[(long)row * ldmod + i + j];
      n[j] = t;
    }
    const float4 o = make_float4(n[0], n[1], n[2], n[3]);
    if (out)
      *reinterpret_cast<float4 *>(orow + i) = o;
    if (sp.hi)
      split_store4(sp, row, i, o);
  }
}

// Same arithmetic (bit for bit: same per-lane summation order, same shuffles) with the row held
// in registers: D == NV * 128, all NV float4 loads of a lane are in flight at once and x is read
// from memory once instead of three dependent passes.  The generic kernel above issues its loads
// one per loop iteration (runtime trip count), which made LayerNorm latency-bound: 12.8 us per
// launch in the batched step, 20 launches per frame.
template <int NV>
__global__ void __launch_bounds__(128)
    layernorm_reg_kernel(const float *x, long ldx, float *out, long ldo, int rows, const float *w, const float *b,
                         float eps, const float *scale, const float *shift, long ldmod, SplitOut sp) {
  PDL_TRIGGER();
  constexpr int D = NV * 128;
  const int lane = threadIdx.x & 31;
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (row >= rows)
    return;
  // the affine weights do not depend on the previous kernel: pull this lane's lines into L1 while it drains
#pragma unroll
  for (int j = 0; j < NV; j++) {
    if (w)
      asm volatile("prefetch.global.L1 [%0];" ::"l"(w + lane * 4 + 128 * j));
    if (b)
      asm volatile("prefetch.global.L1 [%0];" ::"l"(b + lane * 4 + 128 * j));
  }
  PDL_WAIT();
  const float *xr = x + (long)row * ldx;
  float4 v[NV];
#pragma unroll
  for (int j = 0; j < NV; j++)
    v[j] = *reinterpret_cast<const float4 *>(xr + lane * 4 + 128 * j);
  float s = 0.0f;
#pragma unroll
  for (int j = 0; j < NV; j++)
    s += (v[j].x + v[j].y) + (v[j].z + v[j].w);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
    s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mean = s / (float)D;
  float q = 0.0f;
#pragma unroll
  for (int j = 0; j < NV; j++) {
    const float a = v[j].x - mean, c = v[j].y - mean, d = v[j].z - mean, e = v[j].w - mean;
    q += (a * a + c * c) + (d * d + e * e);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
    q += __shfl_xor_sync(0xffffffffu, q, o);
  const float inv = 1.0f / sqrtf(q / (float)D + eps);
  float *orow = out + (long)row * ldo;
#pragma unroll
  for (int j = 0; j < NV; j++) {
    const int i = lane * 4 + 128 * j;
    float n[4] = {(v[j].x - mean) * inv, (v[j].y - mean) * inv, (v[j].z - mean) * inv, (v[j].w - mean) * inv};
    float4 wv = make_float4(1.f, 1.f, 1.f, 1.f), bv = make_float4(0.f, 0.f, 0.f, 0.f);
    if (w)
      wv = __ldg(reinterpret_cast<const float4 *>(w + i));
    if (b)
      bv = __ldg(reinterpret_cast<const float4 *>(b + i));
    const float wa[4] = {wv.x, wv.y, wv.z, wv.w}, ba[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
    for (int k = 0; k < 4; k++) {
      float t = n[k];
      if (w)
        t *= wa[k];
      if (b)
        t += ba[k];
      if (scale)
        t = t * (scale[(long)row * ldmod + i + k] + 1.0f) + shift[(long)row * ldmod + i + k];
      n[k] = t;
    }
    const float4 o = make_float4(n[0], n[1], n[2], n[3]);
    if (out)
      *reinterpret_cast<float4 *>(orow + i) = o;
    if (sp.hi)
      split_store4(sp, row, i, o);
  }
}

static inline bool al16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

int layernorm(const float *x, long ldx, float *out, long ldo, int rows, int D,
              const float *w, const float *b, float eps, const float *scale,
              const float *shift, long ldmod, cudaStream_t st, SplitOut sp) {
  if (rows <= 0)
    return 0;
  static const bool generic_only = getenv("PTTS_LAYERNORM_GENERIC") != nullptr;
  if (!generic_only && (D == 1024 || D == 512 || D == 256) && al16(w) && al16(b)) {
    const dim3 grid((rows + 3) / 4), block(128); // one warp per row, 4 rows per CTA
    if (D == 1024)
      launch_k(layernorm_reg_kernel<8>, grid, block, 0, st, x, ldx, out, ldo, rows, w, b, eps, scale, shift, ldmod, sp);
    else if (D == 512)
      launch_k(layernorm_reg_kernel<4>, grid, block, 0, st, x, ldx, out, ldo, rows, w, b, eps, scale, shift, ldmod, sp);
    else
      launch_k(layernorm_reg_kernel<2>, grid, block, 0, st, x, ldx, out, ldo, rows, w, b, eps, scale, shift, ldmod, sp);
    return 1;
  }
  int ctas = (rows + 7) / 8;
  launch_k(layernorm_kernel, dim3(ctas), dim3(256), 0, st, x, ldx, out, ldo, rows, D, w, b, eps, scale, shift, ldmod, sp);
  return 1;
}

__global__ void __launch_bounds__(256)
    embed_gather_kernel(const int64_t *ids, const float *table, float *out, int D) {
  const int r = blockIdx.x;
  if (ids[r] < 0) // voice-prefix rows are copied in by the host (runtime :104-119)
    return;
  const float4 *src = reinterpret_cast<const float4 *>(table + (long)ids[r] * D);
  float4 *dst = reinterpret_cast<float4 *>(out + (long)r * D);
  for (int i = threadIdx.x; i < D / 4; i += blockDim.x)
    dst[i] = __ldg(src + i);
}

int embed_gather(const int64_t *ids, const float *table, float *out, int rows,
                 int D, cudaStream_t st) {
  if (rows <= 0)
    return 0;
  embed_gather_kernel<<<rows, 256, 0, st>>>(ids, table, out, D);
  return 1;
}

__global__ void nan_to_bos_kernel(const float *frames, const float *bos, float *out,
                                  int n, int ldim, SplitOut sp) {
  PDL_TRIGGER();
  PDL_WAIT();
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n)
    return;
  float v = frames[i];
  v = isnan(v) ? bos[i % ldim] : v;
  out[i] = v;
  if (sp.hi)
    split_store1(sp, i / ldim, i % ldim, v);
}
int nan_to_bos(const float *frames, const float *bos, float *out, int rows,
               int ldim, cudaStream_t st, SplitOut sp) {
  int n = rows * ldim;
  if (n <= 0)
    return 0;
  launch_k(nan_to_bos_kernel, dim3((n + 255) / 256), dim3(256), 0, st, frames, bos, out, n, ldim, sp);
  return 1;
}

__global__ void add_vec_silu_kernel(const float *a, const float *tvec, float *out,
                                    int n, int C, SplitOut sp) {
  PDL_TRIGGER();
  PDL_WAIT();
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n)
    return;
  float y = tvec[i % C] + a[i];
  y = y / (1.0f + expf(-y));
  out[i] = y;
  if (sp.hi)
    split_store1(sp, i / C, i % C, y);
}
int add_vec_silu(const float *a, const float *tvec, float *out, int rows, int C,
                 cudaStream_t st, SplitOut sp) {
  int n = rows * C;
  if (n <= 0)
    return 0;
  launch_k(add_vec_silu_kernel, dim3((n + 255) / 256), dim3(256), 0, st, a, tvec, out, n, C, sp);
  return 1;
}

__global__ void copy_rows_kernel(const float *src, float *dst, int n, int width, SplitOut sp) {
  PDL_TRIGGER();
  PDL_WAIT();
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n)
    return;
  const float v = src[i];
  dst[i] = v;
  if (sp.hi)
    split_store1(sp, i / width, i % width, v);
}
int copy_rows(const float *src, float *dst, int rows, int width, cudaStream_t st, SplitOut sp) {
  int n = rows * width;
  if (n <= 0)
    return 0;
  launch_k(copy_rows_kernel, dim3((n + 255) / 256), dim3(256), 0, st, src, dst, n, width, sp);
  return 1;
}

__global__ void gather_rows_kernel(const float *src, const int64_t *off, float *dst,
                                   int rows, int width, SplitOut sp) {
  PDL_TRIGGER();
  PDL_WAIT();
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * width)
    return;
  int r = i / width, c = i % width;
  const float v = src[off[r] + c];
  dst[i] = v;
  if (sp.hi)
    split_store1(sp, r, c, v);
}
int gather_rows(const float *src, const int64_t *src_off, float *dst, int rows,
                int width, cudaStream_t st, SplitOut sp) {
  int n = rows * width;
  if (n <= 0)
    return 0;
  launch_k(gather_rows_kernel, dim3((n + 255) / 256), dim3(256), 0, st, src, src_off, dst, rows, width, sp);
  return 1;
}
__global__ void scatter_rows_kernel(const float *src, float *dst, const int64_t *off,
                                    int rows, int width) {
  PDL_TRIGGER();
  PDL_WAIT();
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * width)
    return;
  int r = i / width, c = i % width;
  dst[off[r] + c] = src[i];
}
int scatter_rows(const float *src, float *dst, const int64_t *dst_off, int rows,
                 int width, cudaStream_t st) {
  int n = rows * width;
  if (n <= 0)
    return 0;
  launch_k(scatter_rows_kernel, dim3((n + 255) / 256), dim3(256), 0, st, src, dst, dst_off, rows, width);
  return 1;
}

// Depthwise transposed conv: out[p*s + r, c] = in[p, c] w[c, r] + in[p-1, c] w[c, r+s]
// (+ further taps if k > 2s), right-trimmed to s outputs per input position.
__global__ void __launch_bounds__(256)
    upsample_depthwise_kernel(const float *lat, const float *w, float *out, int C,
                              int stride, int k, const int32_t *row_pos,
                              const int32_t *row_frame0) {
  const int row = blockIdx.x;
  const int pos = row_pos[row];
  float *o = out + (long)row * C;
  if (pos < 0) {
    for (int c = threadIdx.x; c < C; c += blockDim.x)
      o[c] = 0.0f;
    return;
  }
  const int f0 = row_frame0[row];
  const int q = pos / stride, r = pos % stride;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float acc = 0.0f;
    // taps in the reference's accumulation order: ascending input index
    for (int kx = r + ((k - 1 - r) / stride) * stride; kx >= 0; kx -= stride) {
      const int ix = q - kx / stride;
      if (ix < 0)
        continue;
      const float iv = lat[(long)(f0 + ix) * C + c];
      acc += iv * __ldg(w + (long)c * k + kx);
    }
    o[c] = acc;
  }
}

int upsample_depthwise(const float *lat, const float *w, float *out, int rows,
                       int C, int stride, int k, const int32_t *row_pos,
                       const int32_t *row_frame0, cudaStream_t st) {
  if (rows <= 0)
    return 0;
  upsample_depthwise_kernel<<<rows, 256, 0, st>>>(lat, w, out, C, stride, k, row_pos,
                                                  row_frame0);
  return 1;
}

// Final Conv1d (Cout = 1, Cin = 64, k <= 4) over channels-last rows.  A CTA owns 256
// consecutive output rows: every input row (plus k-1 rows of left context) is read once
// with fully coalesced float4 loads by a 16-lane group, which forms the row's k partial
// dot products d_t[row] = sum_c w[t][c] * x[row][c]; then pcm[r] = bias + sum_t d_t[r-(k-1)+t].
// (The one-thread-per-sample version read each row k times at a 256-byte lane stride.)
constexpr int FC_ROWS = 256;
__global__ void __launch_bounds__(256)
    final_conv64_kernel(const float *x, int rows, int k, const float *w, float bias,
                        const uint8_t *rowmask, int mask_div, float *pcm) {
  __shared__ float d[4][FC_ROWS + 4];
  const int lane16 = threadIdx.x & 15, grp = threadIdx.x >> 4;
  const long r0 = (long)blockIdx.x * FC_ROWS;
  float4 wt[4];
#pragma unroll
  for (int t = 0; t < 4; t++)
    wt[t] = t < k ? *reinterpret_cast<const float4 *>(w + t * 64 + lane16 * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
  const int span = FC_ROWS + k - 1;
  for (int i = grp; i < span; i += 16) {
    const long r = r0 - (k - 1) + i; // rows before the buffer start are the zeroed slack
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r < rows)
      v = *reinterpret_cast<const float4 *>(x + r * 64 + lane16 * 4);
    float p[4];
#pragma unroll
    for (int t = 0; t < 4; t++) {
      float a = v.x * wt[t].x;
      a = fmaf(v.y, wt[t].y, a);
      a = fmaf(v.z, wt[t].z, a);
      a = fmaf(v.w, wt[t].w, a);
      p[t] = a;
    }
#pragma unroll
    for (int o = 8; o > 0; o >>= 1)
#pragma unroll
      for (int t = 0; t < 4; t++)
        p[t] += __shfl_xor_sync(0xffffffffu, p[t], o);
    if (lane16 == 0) {
#pragma unroll
      for (int t = 0; t < 4; t++)
        d[t][i] = p[t];
    }
  }
  __syncthreads();
  const long r = r0 + threadIdx.x;
  if (r >= rows)
    return;
  float s = bias;
  for (int t = 0; t < k; t++)
    s += d[t][threadIdx.x + t]; // input row r-(k-1)+t sits at index threadIdx.x + t
  if (rowmask && !rowmask[r / mask_div])
    s = 0.0f;
  pcm[r] = s;
}

// Generic fallback: one thread per output sample, weights in smem.
__global__ void __launch_bounds__(256)
    final_conv_kernel(const float *x, int rows, int Cin, int k, const float *w,
                      float bias, const uint8_t *rowmask, int mask_div, float *pcm) {
  extern __shared__ float ws[];
  const int K = Cin * k;
  for (int i = threadIdx.x; i < K; i += blockDim.x)
    ws[i] = w[i];
  __syncthreads();
  const long r = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows)
    return;
  if (rowmask && !rowmask[r / mask_div]) {
    pcm[r] = 0.0f;
    return;
  }
  const float *a = x + (r - (k - 1)) * Cin;
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  for (int i = 0; i < K; i += 4) {
    float4 v = *reinterpret_cast<const float4 *>(a + i);
    s0 = fmaf(v.x, ws[i + 0], s0);
    s1 = fmaf(v.y, ws[i + 1], s1);
    s2 = fmaf(v.z, ws[i + 2], s2);
    s3 = fmaf(v.w, ws[i + 3], s3);
  }
  pcm[r] = ((s0 + s1) + (s2 + s3)) + bias;
}

int final_conv(const float *x, int rows, int Cin, int k, const float *w, float bias,
               const uint8_t *rowmask, int mask_div, float *pcm, cudaStream_t st) {
  if (rows <= 0)
    return 0;
  if (Cin == 64 && k <= 4) {
    final_conv64_kernel<<<(rows + FC_ROWS - 1) / FC_ROWS, 256, 0, st>>>(x, rows, k, w, bias, rowmask, mask_div, pcm);
    return 1;
  }
  final_conv_kernel<<<(rows + 255) / 256, 256, sizeof(float) * Cin * k, st>>>(
      x, rows, Cin, k, w, bias, rowmask, mask_div, pcm);
  return 1;
}

// float PCM -> PCM16 exactly as the reference's streaming WAV writer: the product is formed
// in fp64 (exact for a 24-bit by 15-bit product) and truncated toward zero.
__global__ void pcm_to_s16_kernel(const float *in, int16_t *out, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    double v = (double)in[i];
    v = v != v ? 0.0 : fmax(-1.0, fmin(1.0, v));
    out[i] = (int16_t)(int)(v * 32767.0);
  }
}
int pcm_to_s16(const float *in, int16_t *out, size_t n, cudaStream_t st) {
  if (n == 0)
    return 0;
  size_t blocks = (n + 255) / 256;
  if (blocks > 148 * 16)
    blocks = 148 * 16;
  pcm_to_s16_kernel<<<(unsigned)blocks, 256, 0, st>>>(in, out, n);
  return 1;
}

__global__ void fill_zero_kernel(float4 *p, size_t n4) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n4; i += stride)
    p[i] = make_float4(0.f, 0.f, 0.f, 0.f);
}
int fill_zero(float *p, size_t n, cudaStream_t st) {
  if (n == 0)
    return 0;
  cudaMemsetAsync(p, 0, n * sizeof(float), st);
  return 0;
}

} // namespace ptts
