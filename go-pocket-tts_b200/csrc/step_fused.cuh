// step_fused.cuh — parameters of the single-kernel batch-1 AR frame (step_fused.cu).
#pragma once
#include "kernels.cuh"

namespace ptts {

struct FusedLayer {
  const float *n1w, *n1b, *in_proj, *out_proj, *n2w, *n2b, *l1, *l2;
  float eps1, eps2;
};
struct FusedRes {
  const float *lnw, *lnb, *mlp0w, *mlp0b, *mlp2w, *mlp2b;
  float eps;
};
constexpr int FUSED_MAX_LAYERS = 8;
// default frame kernel when PTTS_FUSED_LL is unset (flag-in-data exchange vs grid barriers)
constexpr bool FUSED_LL_DEFAULT = false;
constexpr bool FUSED_RING_DEFAULT = false; // PTTS_FUSED_RING: the weight-ring frame kernel (step_ring.cu)
constexpr int FUSED_L2_AHEAD_DEFAULT = 0; // PTTS_FUSED_L2_AHEAD: weight phases prefetched into L2 ahead of use
constexpr size_t STEP_FUSED_BARRIER_BYTES = 8192; // top + per-group counters, 128 B apart (<= 63 groups)

// Lives in device memory (one copy per engine); every pointer is a device pointer.
struct FusedParams {
  int D, H, F, L, C, hid, depth, ldim, ada_n;
  FusedLayer layer[FUSED_MAX_LAYERS];
  const float *bos, *in_w, *in_b;  // input_linear [D, ldim]
  const float *on_w, *on_b;        // out_norm
  float on_eps;
  const float *eos_w;
  float eos_b;
  const float *ce_w, *ce_b;        // cond_embed [C, D]
  const float *ada_w, *ada_b;      // all adaLN projections [ada_n, C]
  const float *ip_w, *ip_b;        // input_proj [C, ldim]
  FusedRes rb[FUSED_MAX_LAYERS];
  const float *fin_w, *fin_b;      // final_layer.linear [ldim, C]
  const float *rope_cos, *rope_sin;
  KvGeom kv;
  // activations (the engine's ordinary workspaces, so the debug taps keep working)
  float *x, *qkv, *att, *ff, *fh_in, *fh_last, *fh_eos, *fh_cproj, *fh_ada, *fh_x, *fh_h2, *fh_cur;
  const float *fh_frames, *fh_noise;
  const float *hist, *noise_pool;
  float *hist_w;
  unsigned *barrier;
  // flag-in-data exchange (step_fused_ll_kernel): every cross-CTA activation lives as
  // (value, tag) 8-byte words; `epoch` holds the first tag of the next launch
  uint2 *ll_x, *ll_qkv, *ll_att, *ll_ff, *ll_cproj, *ll_ada, *ll_fx, *ll_h2, *ll_cur;
  unsigned *epoch;
  int poll_ns; // back-off between failed polls (0: none)
  int l2_ahead; // weight phases whose rows are bulk-prefetched into L2 ahead of their use (0: off)
  int *error;
  unsigned long long *dbg; // optional timeline of CTA 0 (globaltimer), tools only
};

size_t step_fused_smem(const FusedParams &hp, int max_keys);
// (value, tag) words the flag-in-data variant needs behind FusedParams::ll_*
size_t step_fused_ll_words(const FusedParams &hp);
void step_fused_ll_assign(FusedParams &hp, uint2 *base);
// CTAs of a cooperative launch (0: the kernel cannot be resident); ll selects the kernel
int step_fused_grid(size_t smem, bool ll);
// returns the number of stream operations enqueued, < 0 on a launch error
int step_fused_launch(const FusedParams *dp, int grid, size_t smem, const int32_t *row_slot, const int32_t *row_pos,
                      const int64_t *off_a, const int64_t *off_b, const int64_t *off_c, int gather_frames,
                      int gather_noise, const float *tvec, int lsd_steps, unsigned *barrier, bool ll, cudaStream_t st);

// The weight-ring variant (step_ring.cu): one CTA per SM, the frame's weights streamed through a shared-memory
// ring by bulk copies.  step_ring_grid returns the CTAs (0: cannot run) and the ring slots / dynamic shared
// memory for an activation area of `act_bytes` (= step_fused_smem).
int step_ring_grid(size_t act_bytes, int *slots, size_t *smem);
int step_ring_launch(const FusedParams *dp, int grid, size_t smem, int slots, const int32_t *row_slot,
                     const int32_t *row_pos, const int64_t *off_a, const int64_t *off_b, const int64_t *off_c,
                     int gather_frames, int gather_noise, const float *tvec, int lsd_steps, unsigned *barrier,
                     cudaStream_t st);

} // namespace ptts
