// engine.cuh — device-resident model + per-sequence state of the native-cuda
// backend (replaces internal/native.Model and FlowLMState of the reference).
#pragma once
#include "../../include/pockettts_cuda.h"
#include "kernels.cuh"
#include "step_fused.cuh"
#include "flow_head.cuh"
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <deque>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

namespace ptts {

void set_error(const char *fmt, ...);
#define PTTS_FAIL(code, ...)                                                   \
  do {                                                                         \
    ::ptts::set_error(__VA_ARGS__);                                            \
    return (code);                                                             \
  } while (0)
#define CUDA_TRY(expr)                                                         \
  do {                                                                         \
    cudaError_t _e = (expr);                                                   \
    if (_e != cudaSuccess) {                                                   \
      ::ptts::set_error("cuda: %s failed: %s", #expr, cudaGetErrorString(_e)); \
      return PTTS_ECUDA;                                                       \
    }                                                                          \
  } while (0)

struct HostTensor {
  std::vector<int64_t> shape;
  std::vector<float> data;
};

struct DLinear {
  float *w = nullptr, *b = nullptr;
  int in = 0, out = 0;
};
struct DLayerNorm {
  float *w = nullptr, *b = nullptr;
  int d = 0;
  float eps = 1e-5f;
};
struct DTLayer {
  DLayerNorm n1, n2;
  DLinear in_proj, out_proj, l1, l2;
  float *ls1 = nullptr, *ls2 = nullptr;
};
// Conv / ConvTranspose in GEMM form: W [N, taps*cin], A rows overlap (lda=cin).
struct DConvG {
  float *w = nullptr, *b = nullptr;
  int cin = 0, cout = 0, taps = 0, stride = 1, N = 0, K = 0;
  void *amap = nullptr; // {cin, taps, rows} TMA map over the input's split planes (taps > 1)
};
struct DResBlk {
  DLayerNorm in_ln;
  DLinear mlp0, mlp2;
};

// bf16 (hi, lo) planes of an activation that feeds a tensor-core GEMM as its A operand,
// written by the producing kernel (split.cuh) and fetched by TMA.
// Layout: [lead slack rows | rows_cap rows] for hi, the same again for lo.
struct ActSplit {
  uint16_t *base = nullptr, *p = nullptr; // allocation / hi plane row 0
  long plane_rows = 0, rows_cap = 0;
  int K = 0, lead = 0;
  void *map = nullptr; // CUtensorMap over {K, rows}: dense A (row stride K)
  SplitOut out() const {
    SplitOut o;
    o.hi = p;
    o.ld = K;
    o.lo_off = plane_rows * (long)K;
    return o;
  }
};

// Shared voice prefix: the K/V of an `audio_prompt` embedding's first
// floor(T/16)*16 positions depend only on the embedding, so they are prefilled once, kept in
// whole KV pages and mapped read-only into the page table of every later sequence with the
// same voice; each sequence prefills only the remaining voice rows and its text.
struct VoiceEntry {
  uint64_t hash = 0;
  int frames = 0;
  std::vector<float> data; // exact copy of the embedding (hash collisions are compared away)
  const float *src = nullptr; // host buffer last confirmed equal to `data`
  std::vector<int> pages;  // shared pages, page_tokens positions each
  int refs = 0;            // live sequences mapping the pages
  uint64_t last_use = 0;
  bool ready = false;      // K/V present (the first sequence's prefill has been issued)
};

struct Seq; // ptts_seq
struct Request;

struct MimiSeg { // one chunk inside a batched Mimi decode
  int n_frames;
  int frame0;    // first row in the gathered latent matrix
  int row0;      // first base row (pad rows included)
  float *host_out; // destination for PCM (may be null)
  int ring = -1;     // stateful streaming delivery: the request's Mimi K/V ring (else -1)
  int abs_frame0 = 0; // ... and the absolute index of the segment's first ROW frame in the request
  int pre = 0;        // latent frames gathered in front of frame0's rows for the upsampling taps only (no rows)
};

struct Engine {
  ptts_config cfg{};
  int device = 0;
  cudaStream_t st = nullptr;
  bool finalized = false;
  std::map<std::string, HostTensor> host;

  // ---- dims ----
  int D = 0, H = 16, Dh = 0, L = 0, F = 0, ldim = 32, nbins = 0;
  int C = 0, depth = 0, nfreq = 0, ada_n = 0;
  int Md = 0, MH = 8, ML = 0, MF = 0, Mctx = 250, up_stride = 16, up_k = 0;
  ptts_model_info info{};

  // ---- device weights ----
  std::vector<void *> allocs;
  float *embed = nullptr, *bos = nullptr;
  std::vector<DTLayer> fl, ml;
  DLinear input_linear, out_eos_lin, cond_embed, input_proj, ada_all, final_linear;
  float out_eos_bias = 0.f;
  DLayerNorm out_norm;
  std::vector<DResBlk> rb;
  float *rope_cos = nullptr, *rope_sin = nullptr;   // [8192, Dh/2]
  float *mrope_cos = nullptr, *mrope_sin = nullptr; // [8192, MDh/2]
  DLinear proj;       // folded latent->mimi projector [Md, ldim] + bias
  float *up_w = nullptr; // depthwise upsample [Md, up_k]
  DConvG init_conv, up[3], res1[3], res2[3];
  float *final_w = nullptr;
  float final_b = 0.f;
  int final_cin = 0, final_k = 0;
  // host copies needed for the (s,t) schedule precompute (flow_net.go:42-83)
  struct TEmb {
    std::vector<float> freqs, w1, b1, w2, b2, alpha;
    int in = 0, hid = 0, out = 0;
  } te[2];
  std::map<int, float *> tvec_cache; // lsd_steps -> device [steps, C]

  // ---- tcgen05 path: split-bf16 copies of the GEMM weights, keyed by the fp32
  //      device pointer the layer code passes as GemmArgs::W ----
  bool tc_on = true;
  int tc_min_m = 9; // M <= 8 stays on the weight-streaming GEMV
  std::unordered_map<const float *, TcWeight *> tc_map;
  std::vector<TcWeight *> tc_owned;
  int make_tc(const float *d_w_fp32, const std::vector<float> &w, int N, int K);
  int debug_gemm(const float *A, int M, int K, const float *W, int N, const float *bias, int act,
                 int path, float *C);
  int debug_gemm_bench(int M, int K, int N, int lda, int path, int reps, float *avg_us);
  int debug_gemm_tma(const float *A, int M, int cin, int taps, const float *W, int N, const float *bias, int act,
                     float *C, float *c_split);

  // split activation planes (A operands of the tcgen05 GEMMs)
  ActSplit s_nrm, s_att, s_ff, s_in, s_last, s_sy, s_cur, s_h, s_h2;
  ActSplit sm_latin, sm_nrm, sm_att, sm_ff, sm_x, sm_y0e, sm_ye[3], sm_he[3];
  // output of a fused residual block (its input planes are still being read as halo rows by
  // neighbouring tiles, so it cannot be written in place) + the next up-conv's map over it
  ActSplit sm_yf[3];
  void *up_alt_amap[3] = {nullptr, nullptr, nullptr};
  std::vector<void *> host_maps;
  int make_act(ActSplit &a, long rows_cap, int K, int lead);
  // conv form of an activation's map: {cin, taps, rows} over the same planes
  int make_conv_map(const ActSplit &a, int cin, int taps, void **out);
  void use_act(GemmArgs &g, const ActSplit &a) const {
    if (tc_on && a.map) {
      g.amap = a.map;
      g.a_lo_row = (int)a.plane_rows;
    }
  }

  // ---- single-kernel batch-1 frame (step_fused.cu) ----
  FusedParams *d_fused = nullptr;
  unsigned *d_fused_bar = nullptr;
  int *d_fused_err = nullptr, *h_fused_err = nullptr;
  unsigned long long *fused_dbg = nullptr;
  int fused_grid = 0;
  size_t fused_smem = 0;
  bool fused_on = false, fused_ll = true;
  int n_bf16_exact = 0; // weights whose lo plane is zero (2-product GEMMs, hi plane only)
  bool fused_ring = false; // PTTS_FUSED_RING: the frame kernel with the shared-memory weight ring (step_ring.cu)
  int ring_grid = 0, ring_slots = 0;
  size_t ring_smem = 0;
  int setup_fused();

  // ---- persistent batched flow head (flow_head.cu): one cooperative kernel per frame ----
  FhParams *d_fh = nullptr;
  unsigned *d_fh_bar = nullptr;
  int *d_fh_err = nullptr, *h_fh_err = nullptr;
  float *fh_cond_bias = nullptr; // cond_embed bias + the (s,t) embedding of the single LSD step
  int fh_grid = 0;
  bool fh_on = false, fh_used = false;
  int setup_flow_head();
  int flow_head_fused(int B);

  // ---- split-K scratch for skinny GEMMs ----
  float *gemm_ws = nullptr;
  long gemm_ws_floats = 0;
  int *gemm_counters = nullptr;
  int gemm_n_counters = 4096;

  // ---- KV pool ----
  void *kv_pool = nullptr; // fp16 elements unless PTTS_ENGINE_KV_F32
  bool kv_half = true;
  int n_pages = 0, page_tokens = 16, max_pages = 0;
  std::vector<int> free_pages, free_slots;
  int32_t *d_pagetab = nullptr;
  std::vector<int32_t> h_pagetab;
  KvGeom kvg{};

  // ---- per-slot frame storage ----
  int max_batch = 0, max_frames = 0;
  float *d_hist = nullptr;  // [slot][max_frames+1][32], row 0 = NaN (BOS)
  float *d_noise = nullptr; // [slot][max_frames][32]

  // ---- FlowLM workspaces (rows <= max_rows) ----
  int max_rows = 0;
  float *x = nullptr, *nrm = nullptr, *qkv = nullptr, *att = nullptr, *ff = nullptr;
  int32_t *d_row_slot = nullptr, *d_row_pos = nullptr;
  int64_t *d_ids = nullptr, *d_off_a = nullptr, *d_off_b = nullptr, *d_off_c = nullptr;
  // pinned host staging
  int32_t *h_row_slot = nullptr, *h_row_pos = nullptr;
  int64_t *h_ids = nullptr, *h_off_a = nullptr, *h_off_b = nullptr, *h_off_c = nullptr;
  float *h_eos = nullptr, *h_lat = nullptr, *h_stage = nullptr;
  // flow head (rows <= max_batch)
  float *fh_in = nullptr, *fh_frames = nullptr, *fh_last = nullptr, *fh_eos = nullptr;
  float *fh_cproj = nullptr, *fh_sy = nullptr, *fh_ada = nullptr, *fh_x = nullptr;
  float *fh_h = nullptr, *fh_h2 = nullptr, *fh_cur = nullptr, *fh_noise = nullptr;

  // ---- Mimi workspaces ----
  int mimi_frames_cap = 0, mimi_rows_cap = 0, mimi_pad = 8;
  float *m_latin = nullptr, *m_lat = nullptr;
  float *m_x = nullptr, *m_nrm = nullptr, *m_qkv = nullptr, *m_att = nullptr, *m_ff = nullptr;
  float *m_y0e = nullptr;
  float *m_y[3] = {nullptr, nullptr, nullptr}, *m_ye[3] = {nullptr, nullptr, nullptr},
        *m_he[3] = {nullptr, nullptr, nullptr};
  float *m_pcm = nullptr;
  int32_t *d_mrow_pos = nullptr, *d_mrow_frame0 = nullptr;
  // stateful streaming decode (f2): per-request K/V rings of the Mimi transformer + per-row descriptors
  float *mimi_ring = nullptr; // [n_stream_states][ML][MIMI_RING][2 Md]
  long ring_state_stride = 0, ring_layer_stride = 0;
  int n_stream_states = 0;
  std::vector<int> free_stream_states;
  int32_t *d_mrow_apos = nullptr;
  int4 *d_mrow_seg = nullptr;
  bool stream_state_on = true;
  uint8_t *d_mmask = nullptr;
  int32_t *h_mrow_pos = nullptr, *h_mrow_frame0 = nullptr;
  uint8_t *h_mmask = nullptr;
  int64_t *d_moff = nullptr, *h_moff = nullptr;

  // ---- asynchronous PCM return (scheduler path): a decode's PCM goes device ->
  //      pinned ring on a copy stream while the next AR steps run; a finisher thread
  //      copies it into the callers' buffers and completes the requests ----
  static constexpr int NPCM = 4;
  static constexpr int NFIN = 3; // finisher threads (first-touch page faults of caller buffers dominate the host copy)
  cudaStream_t st_copy = nullptr;
  // Mimi decode of finished groups runs on its own (low-priority) stream, concurrent with the
  // AR steps on `st`: the skinny AR GEMMs leave SMs idle and wait on each other, the Mimi
  // kernels fill them.  Everything a group needs is per ring slot (latent rows gathered on
  // `st`, row descriptors, PCM); the activation workspace is shared, so groups serialise on
  // st_mimi.  lane_ selects stream + split-K scratch for pgemm / PK.
  cudaStream_t st_mimi = nullptr;
  bool overlap_on = true;
  int lane_ = 0;
  float *gemm_ws_m = nullptr;
  int *gemm_counters_m = nullptr;
  cudaStream_t cur_st() const { return lane_ ? st_mimi : st; }
  float *m_latin_ring[NPCM] = {nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t ev_lat_ready[NPCM]{};
  int32_t *h_mrow_pos_ring[NPCM]{}, *h_mrow_frame0_ring[NPCM]{};
  uint8_t *h_mmask_ring[NPCM]{};
  int64_t *h_moff_ring[NPCM]{};
  int32_t *h_mrow_apos_ring[NPCM]{};
  int4 *h_mrow_seg_ring[NPCM]{};
  float *m_pcm_ring[NPCM] = {nullptr, nullptr, nullptr, nullptr};
  float *h_pcm_ring[NPCM] = {nullptr, nullptr, nullptr, nullptr}, *h_lat_ring[NPCM] = {nullptr, nullptr, nullptr, nullptr};
  int16_t *m_s16_ring[NPCM] = {nullptr, nullptr, nullptr, nullptr}, *h_s16_ring[NPCM] = {nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t ev_dec0[NPCM]{}, ev_pcm_ready[NPCM]{}, ev_copied[NPCM]{}, ev_copy0[NPCM]{};
  struct PcmJob {
    int buf = 0, rc = 0;
    bool copy = false, copy_s16 = false; // some request wants float / PCM16 samples on the host
    std::string msg;
    struct Seg {
      std::shared_ptr<Request> r;
      size_t pcm_off = 0, lat_off = 0;
      int n_frames = 0;   // frames decoded in this segment (window included)
      int skip = 0;       // leading window frames: decoded for context only, not delivered
      int out_frame0 = 0; // position of the first delivered frame in the request's PCM
      int stream_seq = 0; // > 0: k-th delivery of a streaming request (applied in order)
      bool final = true;  // complete the request after this segment
    };
    std::vector<Seg> segs;
  };
  std::mutex fin_mu;
  std::condition_variable fin_cv;
  std::deque<PcmJob> fin_q;
  bool buf_busy[NPCM] = {false, false, false, false};
  std::vector<std::thread> finishers;
  bool fin_running = false, fin_stop = false;
  int pcm_next = 0;
  double fin_mimi_ms = 0, fin_d2h_ms = 0;
  void finisher_loop();
  void ensure_finisher();
  // One span of a request's latent frames to decode: [first, first + count), of which the first
  // `skip` are context only.
  struct DecodeItem {
    std::shared_ptr<Request> r;
    int first = 0, count = 0, skip = 0;
    bool final = true;
    int ring = -1; // stateful delivery: only `skip` (<= 1) frames of conv context are re-decoded
    int pre = 0;   // ... with this many earlier latent frames gathered for the upsampling taps
  };
  int decode_async(const std::vector<DecodeItem> &items);
  int mimi_window_frames() const { // receptive field of one output frame, in latent frames (f2)
    return (ML * (Mctx > 0 ? Mctx - 1 : 8192) + 8 + up_stride - 1) / up_stride + 1;
  }

  // ---- timing / counters ----
  ptts_stage_times times{};
  cudaEvent_t ev[8]{};
  std::map<std::string, std::pair<const float *, size_t>> debug_bufs;

  // ---- per-kernel-class profiling (CUDA events on the launch stream) ----
  struct ProfAcc {
    int64_t launches = 0;
    double ms = 0, flops = 0, bytes = 0;
  };
  struct ProfRec {
    cudaEvent_t a, b;
    std::string cls;
    double flops, bytes;
    std::string cls2;
  };
  bool prof_on = false;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> prof_pool;
  std::vector<ProfRec> prof_pending;
  std::map<std::string, ProfAcc> prof_acc;
  int pgemm(const GemmArgs &g);
  void prof_begin(cudaEvent_t *a, cudaEvent_t *b);
  void prof_end(cudaEvent_t a, cudaEvent_t b, const char *cls, double flops, double bytes,
                const char *cls2 = nullptr);
  void prof_collect(); // after a stream sync

  // ---- scheduler ----
  std::mutex gpu_mu;   // serialises GPU passes (scheduler thread vs stage API)
  std::mutex q_mu;
  std::condition_variable q_cv;
  std::deque<std::shared_ptr<Request>> queue;
  std::unordered_map<int64_t, std::shared_ptr<Request>> tickets;
  int64_t next_ticket = 1;
  std::thread sched;
  bool sched_running = false, sched_stop = false;
  // load seen by a multi-GPU dispatcher (pool.cu): frames still to generate + prompt tokens still to
  // prefill, over every request submitted and not yet completed
  std::atomic<long> inflight_cost{0};
  std::atomic<int> inflight_requests{0};

  ~Engine();
  int create(const ptts_config *c);
  int load_tensor(const char *key, const int64_t *shape, int rank, int dtype, const void *p);
  int finalize();

  // allocation helpers
  int dmalloc(void **p, size_t bytes);
  int upload(const std::vector<float> &v, float **out);
  int upload_key(const std::string &key, float **out, std::vector<int64_t> *shape);

  // voice prefix cache (scheduler path)
  std::vector<std::unique_ptr<VoiceEntry>> voice_cache;
  uint64_t voice_clock = 0;
  int voice_cache_max = 16;
  bool voice_share_on = true;
  int64_t voice_hits = 0, voice_shared_tokens = 0;
  VoiceEntry *voice_lookup(const float *emb, int frames, bool create);
  bool voice_evict_one();
  // sequence resources
  int seq_alloc_shared(int tokens_needed, VoiceEntry *v, Seq **out);
  void voice_donate(Seq *s, VoiceEntry *v);
  int seq_alloc(int tokens_needed, Seq **out);
  void seq_free(Seq *s);
  void release_stream_state(Request &r); // scheduler thread only
  int sync_pagetab(Seq *s);

  // passes (caller holds gpu_mu)
  int flow_forward(int R, int max_keys, bool decode); // decode: one new row per sequence
  int prefill(const std::vector<Seq *> &seqs, const std::vector<const int64_t *> &tok,
              const std::vector<int> &ntok, const std::vector<const float *> &voice,
              const std::vector<int> &nvoice);
  int import_voice_kv(Seq *s, const ptts_voice_kv *vk);
  // CUDA-graph replay of the step body, keyed by (padded rows, lsd steps, input
  // sources); the first use of a key runs eagerly (also configures kernels),
  // the second captures, later ones replay.
  struct StepGraph {
    cudaGraphExec_t exec = nullptr;
    int64_t launches = 0;
    int uses = 0;
  };
  bool graphs_on = true;
  std::map<uint64_t, StepGraph> step_graphs;
  // Small asynchronous Mimi passes (streaming deliveries: a few frames, ~30 launch-bound kernels) replay a
  // captured graph per (rows, frames, segments, ring slot, stateful) instead of launching kernel by kernel.
  std::map<uint64_t, StepGraph> mimi_graphs;
  bool mimi_graphs_on = true; // PTTS_NO_MIMI_GRAPHS
  int step_body(int Bp, int max_keys, int lsd_steps, bool gather_frames, bool gather_noise, bool timed,
                bool capturing = false);
  int step(const std::vector<Seq *> &seqs, const float *frames_in_host,
           const float *noise_host, int lsd_steps, float *lat_out_host,
           float *eos_out_host);
  int flow_head(int B, int lsd_steps);
  int get_tvec(int lsd_steps, float **out);
  // d2h: copy each segment's PCM to its host_out and wait (stage API).  job != null: the
  // asynchronous scheduler path (no wait; PCM lands in the job's pinned ring buffer).
  int mimi_decode(std::vector<MimiSeg> &segs, int total_frames, bool d2h, PcmJob *job = nullptr);
  int decode_seqs(const std::vector<Seq *> &seqs, const std::vector<int> &first,
                  const std::vector<int> &count, const std::vector<float *> &outs,
                  bool d2h);

  // scheduler
  void sched_loop();
  void ensure_sched();
};

struct Seq {
  Engine *e = nullptr;
  int slot = -1;
  VoiceEntry *voice = nullptr; // shared prefix mapped into the first n_shared pages
  int n_shared = 0;
  std::vector<int> pages;
  int kv_len = 0;   // tokens in cache
  int kv_cap = 0;   // tokens reserved
  int frames = 0;   // latent frames generated
};

struct Request {
  int64_t ticket = 0;
  Engine *eng = nullptr; // owner (load accounting)
  long cost = 0;
  std::vector<int64_t> tokens;
  const float *voice_emb = nullptr;
  int voice_frames = 0;
  const ptts_voice_kv *voice_kv = nullptr;
  ptts_gen_cfg cfg{};
  const float *noise = nullptr;
  int noise_rows = 0;
  float *pcm_out = nullptr;
  int64_t pcm_cap = 0;
  int max_steps = 0, lsd_steps = 1;
  // state
  VoiceEntry *voice_entry = nullptr; // cache entry of this request's voice embedding, if any
  Seq *seq = nullptr;
  bool have_cd = false;
  int countdown = 0;
  std::atomic<bool> cancel{false};
  std::vector<float> latents, eos;
  // frame-level streaming (PTTS_REQ_STREAM)
  int stream_frames = 0;  // 0: off
  int stream_issued = 0;  // frames handed to a decode so far
  int stream_seq = 0;     // deliveries issued
  int stream_applied = 0; // deliveries applied by the finishers (in order)
  int frames_ready = 0;   // PCM valid for frames [0, frames_ready)   (guarded by mu)
  int stream_state = -1;  // Mimi K/V ring of a streaming request (-1: bounded re-decode window instead)
  std::chrono::steady_clock::time_point t_submit;
  int finished_at_iter = 0;
  // completion
  std::mutex mu;
  std::condition_variable cv;
  bool done = false;
  ptts_result res{};
  std::string err;
};

int load_safetensors_file(Engine *e, const char *path);
// Publishes a request's outcome (status + per-request message) and wakes its waiter.
void complete_request(const std::shared_ptr<Request> &r, int status, const char *msg);

} // namespace ptts

struct ptts_engine {
  ptts::Engine impl;
};
struct ptts_seq {
  ptts::Seq impl;
};
