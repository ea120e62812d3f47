// kernels.cuh — launch wrappers of the sm_100a kernels behind the native-cuda
// backend.  Layout conventions:
//   * activations are row-major fp32 [rows, channels] ("channels-last" for the
//     SEANet stack, so a causal Conv1d over k taps is a GEMM whose A rows
//     overlap: row stride = Cin, K = k*Cin);
//   * weights are [N, K] row-major (K contiguous), exactly the reference's
//     Linear layout (internal/native/linear.go:12-17);
//   * FlowLM K/V live in a paged pool
//     kv[page][layer][k|v][head][tok_in_page][head_dim].
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ptts {

enum Act : int { ACT_NONE = 0, ACT_GELU = 1, ACT_SILU = 2, ACT_ELU = 3 };

// Optional bf16 (hi, lo) mirror of an fp32 [rows, cols] output: hi plane row stride ld
// (elements), lo plane lo_off elements after the hi plane.  See split.cuh.
struct SplitOut {
  uint16_t *hi = nullptr;
  long ld = 0, lo_off = 0;
};

// Fused GEMM epilogue, applied in this order (matches the reference op order:
// Linear bias -> activation -> LayerScale -> adaLN gate -> residual add):
//   v = acc + bias[n]; v = act(v); v *= colscale[n]; v *= mul[m,n];
//   v *= alpha; v += residual[m,n]; if (!rowmask) v = 0;
//   C[m,n] = v; C2[m,n] = elu(v)
struct GemmEpi {
  const float *bias = nullptr;
  const float *colscale = nullptr;
  const float *mul = nullptr;
  long ldmul = 0;
  const float *residual = nullptr;
  long ldr = 0;
  float alpha = 1.0f;
  int act = ACT_NONE;
  float *C = nullptr;
  long ldc = 0;
  float *C2 = nullptr;
  long ldc2 = 0;
  SplitOut Cs, C2s; // split-bf16 copies of C / C2 for a following tensor-core GEMM
  const uint8_t *rowmask = nullptr; // row m is live iff rowmask[m / mask_div]
  int mask_div = 1;
};

struct GemmArgs {
  const float *A = nullptr; // row m at A + m*lda, K contiguous floats
  long lda = 0;
  const float *W = nullptr; // [N, K]
  long ldw = 0;
  int M = 0, N = 0, K = 0;
  GemmEpi epi;
  // split-K scratch (optional): partial tiles + self-resetting tile counters
  float *ws = nullptr;
  long ws_floats = 0;
  int *counters = nullptr;
  int n_counters = 0;
  // A already split into bf16 planes and described by a TMA map (tc_act_map_init):
  // both operands are then fetched by TMA and A / lda are not read.  a_lo_row = row
  // offset of the lo plane inside the map; a_cin > 0: implicit-GEMM conv over
  // overlapping rows, map dims {cin, taps, rows}, k = tap * cin + c.
  const void *amap = nullptr;
  int a_lo_row = 0, a_cin = 0;
  // optional timeline of CTA 0 (globaltimer ns; tools only)
  unsigned long long *dbg = nullptr;
};

// C = epi(A * W^T).  Chooses the weight-streaming skinny kernel for M <= 8 and
// a tiled SGEMM otherwise.  Requires K % 4 == 0, N % 4 == 0 for the tiled path
// (N == 1 goes through rowdot).  Returns the number of kernels launched.
int gemm(const GemmArgs &g, cudaStream_t st);
// Name of the kernel variant the last gemm() call on this thread selected.
const char *gemm_last_variant();
// ---- tcgen05 path (tcgemm.cu) -------------------------------------------------
// A weight matrix split once into bf16 hi/lo rows [2N, K] plus its TMA
// descriptors (opaque 128-byte CUtensorMap blobs, 64-byte aligned).
struct alignas(64) TcWeight {
  unsigned char map32[128], map64[128], map128[128], map256[128];
  const void *d_split = nullptr;
  int N = 0, K = 0;
  // every weight is exactly a bf16 value (a checkpoint stored as BF16, store.go:339-395): the lo plane is
  // all zero, so the TMA-fed kernels skip it — two products per k-block, half the weight stream
  bool lo_zero = false;
};
int tc_weight_init(TcWeight *w, const void *d_split, int N, int K);
// TMA map (128-byte blob, 64-byte aligned) over split activation planes: `base` points at
// hi row 0, rows_total rows of K bf16 (hi plane then lo plane).  cin > 0 builds the 3-D
// overlapping-row form {cin, taps, rows} for an implicit-GEMM conv (row stride cin).
int tc_act_map_init(void *map, const void *base, long rows_total, int K, int cin, int taps);
bool tc_gemm_supported(const GemmArgs &g);
// Sweep hook: force the tile width / split-K of every following tc_gemm (bn <= 0: automatic).
void tc_force_tile(int bn, int s);
// Same contract as gemm(); W is taken from `w` (g.W is ignored).
int tc_gemm(const TcWeight &w, const GemmArgs &g, cudaStream_t st, const char **variant);

// One SEANet residual block (mimi.go:146-164) as a single kernel (resblock.cu):
//   out = elu(y + b2 + W2 . elu(b1 + W1 . [e(r-2), e(r-1), e(r)])),  e = elu(y) as split planes
// behind `amap` ({C, 3, rows} conv map).  C in {64, 128}, H = C/2.  Returns the number of
// launches, 0 when the shape is not supported (nothing launched).
int resblock_fused(const TcWeight &w1, const TcWeight &w2, const void *amap, int a_lo_row, int M, int C, int H, int taps,
                   const float *b1, const float *b2, const float *y, const uint8_t *rowmask, int mask_div, float *out,
                   SplitOut outs, cudaStream_t st);

// out[m] = dot(A[m,:K], w) + bias0   (N == 1 linear: out_eos)
int rowdot(const float *A, long lda, const float *w, float bias0, int M, int K,
           float *out, cudaStream_t st);

// LayerNorm over the last dim with optional affine and optional adaLN
// modulation: out = (LN(x)*w + b) * (1 + scale[m,:]) + shift[m,:].
// (linear.go:265-329, tensor_util.go:175-193)
int layernorm(const float *x, long ldx, float *out, long ldo, int rows, int D,
              const float *w, const float *b, float eps, const float *scale,
              const float *shift, long ldmod, cudaStream_t st, SplitOut sp = SplitOut());

// out[r,:] = table[ids[r],:]   (conditioner.go:31-53)
int embed_gather(const int64_t *ids, const float *table, float *out, int rows,
                 int D, cudaStream_t st);

struct KvGeom {
  void *pool;           // paged KV pool: __half elements (default) or float (half == 0)
  int half;
  const int32_t *pagetab; // [slots, max_pages]
  int max_pages, page_tokens, n_layers, n_heads, head_dim;
};

// RoPE (interleaved pairs, rope.go:81-105) on q and k of qkv [R, 3*D] in place
// for q, and K/V append into the paged cache at (slot, pos)
// (flow_transformer.go:32-67).
int rope_kv_append(float *qkv, int rows, const int32_t *row_slot,
                   const int32_t *row_pos, const float *cosT, const float *sinT,
                   const KvGeom &kv, int layer, cudaStream_t st);
// Raw K/V rows [T, H*Dh] into pages of one slot (voice model-state import,
// flow_transformer.go:517-631).
int kv_write_rows(const float *rows_k, const float *rows_v, int T, int slot,
                  int pos0, const KvGeom &kv, int layer, cudaStream_t st);
// Reads valid K or V rows of a slot back as [H, T, Dh] (parity tap).
int kv_read_rows(float *out, int T, int slot, const KvGeom &kv, int layer,
                 int which, cudaStream_t st);

// Causal attention of each row against its sequence's paged cache: row r at
// position p sees keys 0..p (attention.go:307-484 with posK[i] = i < len).
// q is read from qkv[r, 0:D], output to out[r, 0:D].
int attention_paged(const float *qkv, long ldq, float *out, long ldo, int rows,
                    const int32_t *row_slot, const int32_t *row_pos,
                    const KvGeom &kv, int layer, int max_keys, cudaStream_t st,
                    SplitOut sp = SplitOut());

// Decode-step fusion of rope_kv_append + attention_paged: every row is the single new token of
// its sequence (qkv is left untouched; K/V of the token are appended to the cache).
int attention_paged_decode(const float *qkv, long ldq, float *out, long ldo, int rows, const int32_t *row_slot,
                           const int32_t *row_pos, const float *cosT, const float *sinT, const KvGeom &kv, int layer,
                           int max_keys, cudaStream_t st, SplitOut sp = SplitOut());
// The same with the K / V rows staged through shared memory by cp.async (attention_decode.cu); 0: not applicable.
int attention_paged_decode_staged(const float *qkv, long ldq, float *out, long ldo, int rows, const int32_t *row_slot,
                           const int32_t *row_pos, const float *cosT, const float *sinT, const KvGeom &kv, int layer,
                           int max_keys, cudaStream_t st, SplitOut sp = SplitOut());

// RoPE in place on q and k of qkv [R, 3*D] with per-row positions (<0 = skip).
int rope_inplace(float *qkv, int rows, int D, int n_heads, const int32_t *row_pos,
                 const float *cosT, const float *sinT, cudaStream_t st);
// Sliding-window causal attention inside contiguous segments: row r at local
// position p = row_pos[r] attends rows r-j, 0 <= j < min(ctx, p+1)
// (mimi.go:386-441, attention.go:473-484).  Pad rows (p < 0) produce zeros.
int attention_window(const float *qkv, long ldq, float *out, long ldo, int rows,
                     int D, int n_heads, const int32_t *row_pos, int ctx,
                     cudaStream_t st, SplitOut sp = SplitOut());

// Stateful streaming decode: the Mimi transformer's K (after RoPE) and V rows of a streaming
// request live in a per-request ring [layer][MIMI_RING rows][K | V] indexed by absolute row % MIMI_RING, so a
// delivery only runs its new rows (plus one frame of conv context) through the decoder: a row attends the
// ring for keys older than its segment and the qkv buffer for keys inside it.
//   row_apos[r] : absolute position of row r in its request (RoPE position), < 0 on pad rows
//   row_seg[r]  : {ring slot of the request or -1, absolute position of the segment's first row,
//                  buffer row of the segment's first row, unused}
constexpr int KV_MAX_SEQ_PAGES = 512; // page ids of one sequence staged in shared memory by the decode attention
constexpr int MIMI_RING = 1024;
constexpr int MIMI_ROPE_ROWS = 8192; // rows of the Mimi transformer's RoPE table
int kv_ring_append(const float *qkv, int rows, int D, const int32_t *row_apos, const int4 *row_seg, float *ring,
                   long state_stride, long layer_off, cudaStream_t st);
int attention_window_stream(const float *qkv, long ldq, float *out, long ldo, int rows, int D, int n_heads,
                            const int32_t *row_apos, const int4 *row_seg, const float *ring, long state_stride,
                            long layer_off, int ctx, cudaStream_t st, SplitOut sp = SplitOut());

// tcgen05 version (attention_tc.cu: flash-style, S and P.V accumulators in TMEM, split-bf16 operands,
// head_dim 64); returns 0 without launching when the shape is not supported.
int attention_window_tc(const float *qkv, long ldq, float *out, long ldo, int rows, int D,
                        int n_heads, const int32_t *row_pos, int ctx, cudaStream_t st,
                        SplitOut sp = SplitOut());
// Legacy tensor-core version (attention_mma.cu: flash-style, split-bf16 mma.sync, head_dim 64);
// returns 0 without launching when the shape is not supported.
int attention_window_mma(const float *qkv, long ldq, float *out, long ldo, int rows, int D,
                         int n_heads, const int32_t *row_pos, int ctx, cudaStream_t st,
                         SplitOut sp = SplitOut());

// x_in[r,:] = isnan(frames[r,i]) ? bos[i] : frames[r,i]  (tensor_util.go:242-271)
int nan_to_bos(const float *frames, const float *bos, float *out, int rows,
               int ldim, cudaStream_t st, SplitOut sp = SplitOut());
// out = silu(a[m,:] + tvec[:])   (flow_net.go:334-343 + resblock SiLU(y))
int add_vec_silu(const float *a, const float *tvec, float *out, int rows, int C,
                 cudaStream_t st, SplitOut sp = SplitOut());
// gathers/scatters of [*,32] latent rows by index
int gather_rows(const float *src, const int64_t *src_off, float *dst, int rows,
                int width, cudaStream_t st, SplitOut sp = SplitOut());
// dst = src over [rows, width] (+ split copy)
int copy_rows(const float *src, float *dst, int rows, int width, cudaStream_t st,
              SplitOut sp = SplitOut());
int scatter_rows(const float *src, float *dst, const int64_t *dst_off, int rows,
                 int width, cudaStream_t st);

// Depthwise ConvTranspose1d (stride s, kernel k<=2s, right trim k-s) from
// per-frame features lat [frames, C] into the padded channels-last layout:
// out[row, c], row_frame[row] = source frame index of that row's segment start
// (or -1 on pad rows), row_pos = local position.  (mimi.go:116-125,
// convtranspose1d.go:154-202)
int upsample_depthwise(const float *lat, const float *w, float *out, int rows,
                       int C, int stride, int k, const int32_t *row_pos,
                       const int32_t *row_frame0, cudaStream_t st);
// Final Conv1d Cin->1 over channels-last rows: pcm[r] = dot(x[r-k+1..r, :], w) + b
int final_conv(const float *x, int rows, int Cin, int k, const float *w,
               float bias, const uint8_t *rowmask, int mask_div, float *pcm,
               cudaStream_t st);
int fill_zero(float *p, size_t n, cudaStream_t st);
// out[i] = int16(clamp(in[i], -1, 1) * 32767) truncated toward zero (wav_stream.go:44-54)
int pcm_to_s16(const float *in, int16_t *out, size_t n, cudaStream_t st);

} // namespace ptts
