// engine.cu — weight upload, KV pool and the three device passes (prefill, AR
// step + flow head, Mimi decode) of the native-cuda backend.
#include "engine.cuh"
#include "launch.cuh"
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace ptts {

static thread_local char g_err[512] = "";
void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof g_err, fmt, ap);
  va_end(ap);
}
const char *last_error() { return g_err; }

// ---------------------------------------------------------------------------
// create / destroy
// ---------------------------------------------------------------------------
int Engine::create(const ptts_config *c) {
  cfg = *c;
  if (cfg.max_batch <= 0) cfg.max_batch = 256;
  if (cfg.page_tokens <= 0) {
    const char *pt = getenv("PTTS_PAGE_TOKENS"); // experiments: K / V rows of one head are contiguous per page
    cfg.page_tokens = pt && atoi(pt) > 0 ? atoi(pt) : 16;
  }
  if (cfg.max_seq_tokens <= 0) cfg.max_seq_tokens = 1024;
  if (cfg.max_seq_tokens > 8192)
    PTTS_FAIL(PTTS_EINVAL, "native-cuda: max_seq_tokens %d exceeds the RoPE table (8192)",
              cfg.max_seq_tokens);
  if (cfg.max_frames <= 0) cfg.max_frames = 512;
  if (cfg.mimi_group_frames <= 0) cfg.mimi_group_frames = 1536;
  if (cfg.max_prefill_tokens <= 0) cfg.max_prefill_tokens = 8192;
  if (cfg.max_prefill_tokens < cfg.max_seq_tokens) cfg.max_prefill_tokens = cfg.max_seq_tokens;
  int ndev = 0;
  cudaError_t ce = cudaGetDeviceCount(&ndev);
  if (ce != cudaSuccess || ndev == 0)
    PTTS_FAIL(PTTS_ECUDA, "native-cuda runtime unavailable: no CUDA device (%s)",
              cudaGetErrorString(ce));
  if (cfg.device < 0 || cfg.device >= ndev)
    PTTS_FAIL(PTTS_EINVAL, "native-cuda: device %d out of range [0,%d)", cfg.device, ndev);
  device = cfg.device;
  tc_on = !(cfg.flags & PTTS_ENGINE_NO_TENSOR_CORES) && !getenv("PTTS_NO_TC");
  graphs_on = !(cfg.flags & PTTS_ENGINE_NO_GRAPHS) && !getenv("PTTS_NO_GRAPHS");
  CUDA_TRY(cudaSetDevice(device));
  cudaDeviceProp prop;
  CUDA_TRY(cudaGetDeviceProperties(&prop, device));
  if (prop.major < 10)
    PTTS_FAIL(PTTS_ECUDA, "native-cuda runtime unavailable: device %d is sm_%d%d, need sm_100",
              device, prop.major, prop.minor);
  CUDA_TRY(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
  for (auto &e : ev)
    CUDA_TRY(cudaEventCreate(&e));
  return PTTS_OK;
}

Engine::~Engine() {
  {
    std::unique_lock<std::mutex> lk(q_mu);
    sched_stop = true;
  }
  q_cv.notify_all();
  if (sched.joinable())
    sched.join();
  {
    std::unique_lock<std::mutex> lk(fin_mu);
    fin_stop = true;
  }
  fin_cv.notify_all();
  for (auto &t : finishers)
    if (t.joinable())
      t.join();
  if (st) {
    cudaSetDevice(device);
    cudaStreamSynchronize(st);
  }
  for (void *p : allocs)
    cudaFree(p);
  for (TcWeight *t : tc_owned)
    delete t;
  for (void *m : host_maps)
    free(m);
  for (auto &kv : step_graphs)
    if (kv.second.exec)
      cudaGraphExecDestroy(kv.second.exec);
  for (auto &kv : mimi_graphs)
    if (kv.second.exec)
      cudaGraphExecDestroy(kv.second.exec);
  auto fh = [](void *p) {
    if (p)
      cudaFreeHost(p);
  };
  fh(h_row_slot); fh(h_row_pos); fh(h_ids); fh(h_off_a); fh(h_off_b); fh(h_off_c);
  fh(h_fused_err); fh(h_fh_err);
  fh(h_eos); fh(h_lat); fh(h_stage); fh(h_mrow_pos); fh(h_mrow_frame0); fh(h_mmask); fh(h_moff);
  for (int i = 0; i < NPCM; i++) {
    fh(h_pcm_ring[i]);
    fh(h_lat_ring[i]);
    fh(h_s16_ring[i]);
    fh(h_mrow_pos_ring[i]); fh(h_mrow_frame0_ring[i]); fh(h_mmask_ring[i]); fh(h_moff_ring[i]);
    fh(h_mrow_apos_ring[i]); fh(h_mrow_seg_ring[i]);
    for (cudaEvent_t e2 : {ev_dec0[i], ev_pcm_ready[i], ev_copied[i], ev_copy0[i], ev_lat_ready[i]})
      if (e2)
        cudaEventDestroy(e2);
  }
  if (st_mimi) {
    cudaStreamSynchronize(st_mimi);
    cudaStreamDestroy(st_mimi);
  }
  if (st_copy)
    cudaStreamDestroy(st_copy);
  for (auto &e : ev)
    if (e)
      cudaEventDestroy(e);
  if (st)
    cudaStreamDestroy(st);
}

int Engine::dmalloc(void **p, size_t bytes) {
  if (bytes == 0)
    bytes = 256;
  CUDA_TRY(cudaMalloc(p, bytes));
  allocs.push_back(*p);
  return PTTS_OK;
}

int Engine::upload(const std::vector<float> &v, float **out) {
  int rc = dmalloc((void **)out, v.size() * sizeof(float));
  if (rc)
    return rc;
  CUDA_TRY(cudaMemcpy(*out, v.data(), v.size() * sizeof(float), cudaMemcpyHostToDevice));
  return PTTS_OK;
}

int Engine::upload_key(const std::string &key, float **out, std::vector<int64_t> *shape) {
  auto it = host.find(key);
  if (it == host.end())
    PTTS_FAIL(PTTS_EINVAL, "native: tensor %s not found", key.c_str());
  if (shape)
    *shape = it->second.shape;
  return upload(it->second.data, out);
}

// ---------------------------------------------------------------------------
// profiling: one CUDA-event pair per profiled launch, resolved after the next
// stream sync.  Algorithmic FLOPs/bytes are recorded with the launch so the
// bench can quote achieved TFLOP/s / GB/s per kernel class.
// ---------------------------------------------------------------------------
void Engine::prof_begin(cudaEvent_t *a, cudaEvent_t *b) {
  if (prof_pool.empty()) {
    cudaEventCreate(a);
    cudaEventCreate(b);
  } else {
    *a = prof_pool.back().first;
    *b = prof_pool.back().second;
    prof_pool.pop_back();
  }
  cudaEventRecord(*a, cur_st());
}
void Engine::prof_end(cudaEvent_t a, cudaEvent_t b, const char *cls, double flops, double bytes,
                      const char *cls2) {
  cudaEventRecord(b, cur_st());
  prof_pending.push_back(ProfRec{a, b, cls, flops, bytes, cls2 ? cls2 : ""});
}
void Engine::prof_collect() {
  for (auto &r : prof_pending) {
    float ms = 0;
    if (cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) {
      for (const std::string *c : {&r.cls, &r.cls2}) {
        if (c->empty())
          continue;
        ProfAcc &acc = prof_acc[*c];
        acc.launches++;
        acc.ms += ms;
        acc.flops += r.flops;
        acc.bytes += r.bytes;
      }
    }
    prof_pool.push_back({r.a, r.b});
  }
  prof_pending.clear();
}
// Launches `call` (returns the launch count), timed as class `cls` when profiling.
#define PK(cls, flops, bytes, call)                                            \
  do {                                                                         \
    if (prof_on) {                                                             \
      cudaEvent_t _a, _b;                                                      \
      prof_begin(&_a, &_b);                                                    \
      nl += (call);                                                            \
      prof_end(_a, _b, cls, (double)(flops), (double)(bytes));                 \
    } else                                                                     \
      nl += (call);                                                            \
  } while (0)
static inline uint16_t f32_to_bf16_rn(float x) {
  uint32_t u;
  memcpy(&u, &x, 4);
  u = (u + 0x7FFFu + ((u >> 16) & 1u)) >> 16;
  return (uint16_t)u;
}
static inline float bf16_to_f32(uint16_t h) {
  uint32_t u = (uint32_t)h << 16;
  float x;
  memcpy(&x, &u, 4);
  return x;
}

// Splits W [N,K] fp32 into bf16 hi rows [0,N) and lo rows [N,2N) (x ~= hi + lo,
// ~16 mantissa bits), uploads it and builds the TMA descriptors.
int Engine::make_tc(const float *d_w_fp32, const std::vector<float> &w, int N, int K) {
  if (!tc_on || K % 8 != 0 || N % 16 != 0 || K < 32 || N < 32)
    return PTTS_OK; // stays on the CUDA-core path
  std::vector<uint16_t> sp((size_t)2 * N * K);
  bool lo_zero = true;
  for (size_t i = 0; i < (size_t)N * K; i++) {
    uint16_t h = f32_to_bf16_rn(w[i]);
    sp[i] = h;
    const uint16_t lo = f32_to_bf16_rn(w[i] - bf16_to_f32(h));
    sp[(size_t)N * K + i] = lo;
    lo_zero = lo_zero && (lo & 0x7fffu) == 0;
  }
  const bool no_fast = getenv("PTTS_NO_BF16_FAST") != nullptr;
  void *d = nullptr;
  int rc = dmalloc(&d, sp.size() * 2);
  if (rc)
    return rc;
  CUDA_TRY(cudaMemcpy(d, sp.data(), sp.size() * 2, cudaMemcpyHostToDevice));
  TcWeight *t = new TcWeight();
  tc_owned.push_back(t);
  int r = tc_weight_init(t, d, N, K);
  if (r)
    PTTS_FAIL(PTTS_ECUDA, "native-cuda: cuTensorMapEncodeTiled failed (%d) for a [%d,%d] weight", r, N, K);
  t->lo_zero = lo_zero && !no_fast;
  if (t->lo_zero)
    n_bf16_exact++;
  tc_map[d_w_fp32] = t;
  return PTTS_OK;
}

int Engine::make_act(ActSplit &a, long rows_cap, int K, int lead) {
  a.rows_cap = rows_cap;
  a.K = K;
  a.lead = lead;
  a.plane_rows = rows_cap + lead;
  const size_t elems = (size_t)2 * a.plane_rows * K;
  int rc = dmalloc((void **)&a.base, elems * 2);
  if (rc)
    return rc;
  CUDA_TRY(cudaMemset(a.base, 0, elems * 2));
  a.p = a.base + (size_t)lead * K;
  if (!tc_on)
    return PTTS_OK;
  a.map = aligned_alloc(64, 128);
  host_maps.push_back(a.map);
  if (tc_act_map_init(a.map, a.p, a.plane_rows + rows_cap, K, 0, 0))
    PTTS_FAIL(PTTS_ECUDA, "native-cuda: cuTensorMapEncodeTiled failed for a [%ld,%d] activation", rows_cap, K);
  return PTTS_OK;
}

int Engine::make_conv_map(const ActSplit &a, int cin, int taps, void **out) {
  *out = nullptr;
  if (!tc_on)
    return PTTS_OK;
  if (a.lead < taps - 1 || cin != a.K)
    PTTS_FAIL(PTTS_EINVAL, "native-cuda: conv map needs %d slack rows of %d channels", taps - 1, cin);
  void *m = aligned_alloc(64, 128);
  host_maps.push_back(m);
  // row m, tap t -> plane row m - (taps-1) + t
  if (tc_act_map_init(m, a.p - (size_t)(taps - 1) * cin, a.plane_rows + a.rows_cap + taps - 1, cin * taps, cin, taps))
    PTTS_FAIL(PTTS_ECUDA, "native-cuda: cuTensorMapEncodeTiled failed for a conv map (cin %d, taps %d)", cin, taps);
  *out = m;
  return PTTS_OK;
}

int Engine::pgemm(const GemmArgs &g0) {
  GemmArgs g = g0;
  cudaStream_t st = cur_st(); // lane 1 (Mimi decode) has its own stream and split-K scratch
  g.ws = lane_ ? gemm_ws_m : gemm_ws;
  g.ws_floats = gemm_ws_floats;
  g.counters = lane_ ? gemm_counters_m : gemm_counters;
  g.n_counters = gemm_n_counters;
  const TcWeight *tw = nullptr;
  if (tc_on && g.M >= tc_min_m && tc_gemm_supported(g)) {
    auto it = tc_map.find(g.W);
    if (it != tc_map.end())
      tw = it->second;
  }
  if (!tw) {
    g.amap = nullptr; // CUDA-core paths read the fp32 activation
    if (!g.A) {
      set_error("native-cuda: GEMM [%d,%d]x[%d] has no fp32 A and no tensor-core path", g.M, g.K, g.N);
      return 0;
    }
  }
  const char *variant = "";
  cudaEvent_t a, b;
  if (prof_on)
    prof_begin(&a, &b);
  int n;
  if (tw) {
    n = tc_gemm(*tw, g, st, &variant);
    if (n < 0) { // could not configure the kernel: fall back to the CUDA-core GEMM
      n = gemm(g, st);
      variant = gemm_last_variant();
    }
  } else {
    n = gemm(g, st);
    variant = gemm_last_variant();
  }
  if (prof_on) {
    double kk = g.K < g.lda ? g.K : g.lda; // overlapping conv rows: unique A bytes
    // second key: per weight shape ("shape/..." entries are a breakdown, not classes)
    char shp[96];
    snprintf(shp, sizeof shp, "shape/N%d_K%d_lda%ld_%s", g.N, g.K, g.lda, g.M <= 1024 ? "skinny" : "wide");
    prof_end(a, b, variant, 2.0 * g.M * g.N * g.K,
             4.0 * ((double)g.N * g.K + (double)g.M * kk + (double)g.M * g.N), shp);
  }
  return n;
}

// Stand-alone GEMM through a chosen path (0 auto, 1 CUDA cores, 2 tcgen05):
// unit-test hook for the kernels.
int Engine::debug_gemm(const float *A, int M, int K, const float *W, int N, const float *bias, int act,
                       int path, float *C) {
  float *dA = nullptr, *dW = nullptr, *dB = nullptr, *dC = nullptr;
  CUDA_TRY(cudaMalloc(&dA, (size_t)M * K * 4));
  CUDA_TRY(cudaMalloc(&dW, (size_t)N * K * 4));
  CUDA_TRY(cudaMalloc(&dC, (size_t)M * N * 4));
  CUDA_TRY(cudaMemcpy(dA, A, (size_t)M * K * 4, cudaMemcpyHostToDevice));
  CUDA_TRY(cudaMemcpy(dW, W, (size_t)N * K * 4, cudaMemcpyHostToDevice));
  if (bias) {
    CUDA_TRY(cudaMalloc(&dB, (size_t)N * 4));
    CUDA_TRY(cudaMemcpy(dB, bias, (size_t)N * 4, cudaMemcpyHostToDevice));
  }
  GemmArgs g;
  g.A = dA; g.lda = K; g.W = dW; g.ldw = K; g.M = M; g.N = N; g.K = K;
  g.epi.bias = dB; g.epi.act = act; g.epi.C = dC; g.epi.ldc = N;
  int rc = PTTS_OK;
  const bool saved = tc_on;
  const int saved_min = tc_min_m;
  if (path == 2) {
    tc_on = true;
    tc_min_m = 1;
    std::vector<float> hw(W, W + (size_t)N * K);
    rc = make_tc(dW, hw, N, K);
    if (rc == PTTS_OK && (!tc_map.count(dW) || !tc_gemm_supported(g))) {
      set_error("native-cuda: debug_gemm: shape [%d,%d]x[%d,%d] not supported by the tcgen05 path", M, K, N, K);
      rc = PTTS_EINVAL;
    }
  } else if (path == 1) {
    tc_on = false;
  }
  if (rc == PTTS_OK) {
    times.kernel_launches += pgemm(g);
    cudaError_t e = cudaStreamSynchronize(st);
    if (e == cudaSuccess)
      e = cudaGetLastError();
    if (e != cudaSuccess) {
      set_error("cuda: debug_gemm failed: %s", cudaGetErrorString(e));
      rc = PTTS_ECUDA;
    } else {
      cudaMemcpy(C, dC, (size_t)M * N * 4, cudaMemcpyDeviceToHost);
    }
  }
  tc_on = saved;
  tc_min_m = saved_min;
  tc_map.erase(dW);
  cudaFree(dA); cudaFree(dW); cudaFree(dC);
  if (dB) cudaFree(dB);
  return rc;
}

// TMA-fed GEMM unit-test hook: A [M, cin] is split into bf16 planes on the device and
// fetched by TMA; taps > 1 runs the implicit-GEMM conv form (W is [N, taps*cin], row m
// sees rows m-taps+1 .. m, rows < 0 are the zero slack).  C (fp32) and, if c_split is
// not null, the split-plane copy of the same output (hi + lo, reconstructed on the host).
int Engine::debug_gemm_tma(const float *A, int M, int cin, int taps, const float *W, int N, const float *bias,
                           int act, float *C, float *c_split) {
  if (!tc_on)
    PTTS_FAIL(PTTS_ESTATE, "native-cuda: tensor-core path is disabled");
  const int K = cin * taps;
  float *dA = nullptr, *dW = nullptr, *dB = nullptr, *dC = nullptr;
  CUDA_TRY(cudaMalloc(&dA, (size_t)M * cin * 4));
  CUDA_TRY(cudaMalloc(&dW, (size_t)N * K * 4));
  CUDA_TRY(cudaMalloc(&dC, (size_t)M * N * 4));
  CUDA_TRY(cudaMemcpy(dA, A, (size_t)M * cin * 4, cudaMemcpyHostToDevice));
  CUDA_TRY(cudaMemcpy(dW, W, (size_t)N * K * 4, cudaMemcpyHostToDevice));
  if (bias) {
    CUDA_TRY(cudaMalloc(&dB, (size_t)N * 4));
    CUDA_TRY(cudaMemcpy(dB, bias, (size_t)N * 4, cudaMemcpyHostToDevice));
  }
  ActSplit in, out;
  int rc = make_act(in, M, cin, 8);
  if (rc == PTTS_OK)
    rc = make_act(out, M, N, 0);
  std::vector<float> hw(W, W + (size_t)N * K);
  const int saved_min = tc_min_m;
  tc_min_m = 1;
  if (rc == PTTS_OK)
    rc = make_tc(dW, hw, N, K);
  GemmArgs g;
  g.W = dW; g.ldw = K; g.M = M; g.N = N; g.K = K; g.lda = cin;
  g.epi.bias = dB; g.epi.act = act; g.epi.C = dC; g.epi.ldc = N;
  if (c_split)
    g.epi.Cs = out.out();
  if (rc == PTTS_OK) {
    copy_rows(dA, dA, M, cin, st, in.out());
    if (taps > 1) {
      void *cm = nullptr;
      rc = make_conv_map(in, cin, taps, &cm);
      g.amap = cm;
      g.a_lo_row = (int)in.plane_rows;
      g.a_cin = cin;
    } else
      use_act(g, in);
  }
  if (rc == PTTS_OK && (!tc_map.count(dW) || !tc_gemm_supported(g))) {
    set_error("native-cuda: debug_gemm_tma: shape [%d,%d]x[%d] not supported by the tcgen05 path", M, K, N);
    rc = PTTS_EINVAL;
  }
  if (rc == PTTS_OK) {
    times.kernel_launches += pgemm(g);
    cudaError_t e = cudaStreamSynchronize(st);
    if (e == cudaSuccess)
      e = cudaGetLastError();
    if (e != cudaSuccess) {
      set_error("cuda: debug_gemm_tma failed: %s", cudaGetErrorString(e));
      rc = PTTS_ECUDA;
    } else {
      cudaMemcpy(C, dC, (size_t)M * N * 4, cudaMemcpyDeviceToHost);
      if (c_split) {
        std::vector<uint16_t> hi((size_t)M * N), lo((size_t)M * N);
        cudaMemcpy(hi.data(), out.p, hi.size() * 2, cudaMemcpyDeviceToHost);
        cudaMemcpy(lo.data(), out.p + out.plane_rows * (long)N, lo.size() * 2, cudaMemcpyDeviceToHost);
        for (size_t i2 = 0; i2 < hi.size(); i2++)
          c_split[i2] = bf16_to_f32(hi[i2]) + bf16_to_f32(lo[i2]);
      }
    }
  }
  tc_min_m = saved_min;
  tc_map.erase(dW);
  cudaFree(dA); cudaFree(dW); cudaFree(dC);
  if (dB) cudaFree(dB);
  return rc;
}

// Times one GEMM shape on device-resident data (CUDA events around `reps`
// back-to-back launches).  lda < K exercises the overlapping-row conv form.
int Engine::debug_gemm_bench(int M, int K, int N, int lda, int path, int reps, float *avg_us) {
  // path 3 = tcgen05 with the weight rotated over enough copies to defeat the L2
  // (the AR step streams 341 MB of weights per frame, nothing stays cached)
  if (lda <= 0)
    lda = K;
  // path 4 = path 3 with A pre-split into bf16 planes and fetched by TMA
  const bool atma = path == 4;
  const bool rotate = path == 3 || atma;
  if (rotate)
    path = 2;
  int copies = 1;
  if (rotate) {
    copies = (int)((size_t)200e6 / ((size_t)N * K * 4)) + 1;
    if (copies > 24) copies = 24;
  }
  float *dA = nullptr, *dC = nullptr;
  std::vector<float *> dW(copies, nullptr);
  const size_t a_floats = (size_t)M * lda + K + 64;
  CUDA_TRY(cudaMalloc(&dA, a_floats * 4));
  for (auto &w : dW)
    CUDA_TRY(cudaMalloc(&w, (size_t)N * K * 4));
  CUDA_TRY(cudaMalloc(&dC, (size_t)M * N * 4));
  std::vector<float> hw((size_t)N * K), ha(a_floats);
  for (size_t i = 0; i < hw.size(); i++)
    hw[i] = (float)((int)(i * 2654435761u >> 20) % 2001 - 1000) * 1e-3f;
  for (size_t i = 0; i < ha.size(); i++)
    ha[i] = (float)((int)(i * 40503u >> 7) % 2001 - 1000) * 1e-3f;
  for (auto &w : dW)
    CUDA_TRY(cudaMemcpy(w, hw.data(), hw.size() * 4, cudaMemcpyHostToDevice));
  CUDA_TRY(cudaMemcpy(dA, ha.data(), ha.size() * 4, cudaMemcpyHostToDevice));
  GemmArgs g;
  g.A = dA; g.lda = lda; g.ldw = K; g.M = M; g.N = N; g.K = K;
  g.epi.C = dC; g.epi.ldc = N;
  const bool saved = tc_on, saved_prof = prof_on;
  const int saved_min = tc_min_m;
  prof_on = false;
  int rc = PTTS_OK;
  if (path == 2) {
    tc_on = true;
    tc_min_m = 1;
    for (auto &w : dW)
      if (rc == PTTS_OK)
        rc = make_tc(w, hw, N, K);
  } else if (path == 1) {
    tc_on = false;
  }
  ActSplit tmp;
  float *dA2 = nullptr;
  if (rc == PTTS_OK && atma && lda == K) {
    rc = make_act(tmp, M, K, 0);
    if (rc == PTTS_OK) {
      CUDA_TRY(cudaMalloc(&dA2, (size_t)M * K * 4));
      copy_rows(dA, dA2, M, K, st, tmp.out());
      use_act(g, tmp);
    }
  } else if (rc == PTTS_OK && atma && K % lda == 0) { // conv form: taps = K / lda overlapping rows
    const int taps = K / lda;
    rc = make_act(tmp, M, lda, 8);
    void *cm = nullptr;
    if (rc == PTTS_OK)
      rc = make_conv_map(tmp, lda, taps, &cm);
    if (rc == PTTS_OK) {
      CUDA_TRY(cudaMalloc(&dA2, (size_t)M * lda * 4));
      copy_rows(dA, dA2, M, lda, st, tmp.out());
      g.amap = cm;
      g.a_lo_row = (int)tmp.plane_rows;
      g.a_cin = lda;
    }
  }
  if (rc == PTTS_OK) {
    for (int i = 0; i < 3; i++) {
      g.W = dW[i % copies];
      pgemm(g);
    }
    cudaEventRecord(ev[0], st);
    for (int i = 0; i < reps; i++) {
      g.W = dW[i % copies];
      pgemm(g);
    }
    cudaEventRecord(ev[1], st);
    cudaError_t e = cudaStreamSynchronize(st);
    if (e == cudaSuccess)
      e = cudaGetLastError();
    if (e != cudaSuccess) {
      set_error("cuda: debug_gemm_bench failed: %s", cudaGetErrorString(e));
      rc = PTTS_ECUDA;
    } else {
      float ms = 0;
      cudaEventElapsedTime(&ms, ev[0], ev[1]);
      *avg_us = 1e3f * ms / (float)reps;
    }
    if (getenv("PTTS_GEMM_TIMELINE")) { // CTA 0's timeline of one more launch (tools only)
      unsigned long long *dd = nullptr, hd[128];
      cudaMalloc(&dd, sizeof hd);
      cudaMemset(dd, 0, sizeof hd);
      g.W = dW[reps % copies];
      g.dbg = dd;
      pgemm(g);
      g.dbg = nullptr;
      cudaStreamSynchronize(st);
      cudaMemcpy(hd, dd, sizeof hd, cudaMemcpyDeviceToHost);
      cudaFree(dd);
      fprintf(stderr, "timeline M=%d K=%d N=%d: prologue %lld", M, K, N, (long long)(hd[1] - hd[0]));
      const char *lbl[3] = {"mma-ready", "A-written / item-epilogue-done", "tma-issue"};
      const int base[3] = {2, 48, 96};
      for (int t = 0; t < 3; t++) {
        fprintf(stderr, "\n  %s:", lbl[t]);
        for (int i = 0; i < 32; i++)
          if (hd[base[t] + i])
            fprintf(stderr, " %lld", (long long)(hd[base[t] + i] - hd[0]));
      }
      fprintf(stderr, "\n  mma-all-issued %lld  epi-acc-ready %lld  epi-tile-done %lld  epi-exit %lld  cta-sync %lld\n",
              (long long)(hd[40] - hd[0]), (long long)(hd[41] - hd[0]), (long long)(hd[42] - hd[0]),
              (long long)(hd[43] - hd[0]), (long long)(hd[44] - hd[0]));
    }
  }
  tc_on = saved;
  tc_min_m = saved_min;
  prof_on = saved_prof;
  for (auto &w : dW) {
    tc_map.erase(w);
    cudaFree(w);
  }
  cudaFree(dA); cudaFree(dC);
  if (dA2) cudaFree(dA2);
  return rc;
}

// ---------------------------------------------------------------------------
// weights
// ---------------------------------------------------------------------------
static inline float half_to_float(uint16_t h) {
  uint32_t sign = (uint32_t)(h & 0x8000) << 16;
  uint32_t exp = (h >> 10) & 0x1f, man = h & 0x3ff;
  uint32_t f;
  if (exp == 0) {
    if (man == 0) {
      f = sign;
    } else {
      exp = 127 - 15 + 1;
      while (!(man & 0x400)) {
        man <<= 1;
        exp--;
      }
      man &= 0x3ff;
      f = sign | (exp << 23) | (man << 13);
    }
  } else if (exp == 31) {
    f = sign | 0x7f800000u | (man << 13);
  } else {
    f = sign | ((exp + 127 - 15) << 23) | (man << 13);
  }
  float r;
  memcpy(&r, &f, 4);
  return r;
}

// internal/safetensors/store.go:339-395 (decodeTensorData): F32 copied, F16 and
// BF16 widened to fp32.
int Engine::load_tensor(const char *key, const int64_t *shape, int rank, int dtype,
                        const void *p) {
  if (finalized)
    PTTS_FAIL(PTTS_ESTATE, "native-cuda: weights already finalized");
  if (!key || !p || rank < 0 || rank > 6)
    PTTS_FAIL(PTTS_EINVAL, "native-cuda: load_tensor: bad arguments");
  HostTensor t;
  size_t n = 1;
  for (int i = 0; i < rank; i++) {
    if (shape[i] < 0)
      PTTS_FAIL(PTTS_EINVAL, "native-cuda: load_tensor %s: negative dim", key);
    t.shape.push_back(shape[i]);
    n *= (size_t)shape[i];
  }
  t.data.resize(n);
  if (dtype == PTTS_DTYPE_F32) {
    memcpy(t.data.data(), p, n * 4);
  } else if (dtype == PTTS_DTYPE_BF16) {
    const uint16_t *s = (const uint16_t *)p;
    for (size_t i = 0; i < n; i++) {
      uint32_t u = (uint32_t)s[i] << 16;
      memcpy(&t.data[i], &u, 4);
    }
  } else if (dtype == PTTS_DTYPE_F16) {
    const uint16_t *s = (const uint16_t *)p;
    for (size_t i = 0; i < n; i++)
      t.data[i] = half_to_float(s[i]);
  } else {
    PTTS_FAIL(PTTS_EINVAL, "native-cuda: load_tensor %s: unsupported dtype %d", key, dtype);
  }
  host[key] = std::move(t);
  return PTTS_OK;
}

namespace {
struct Loader {
  Engine *e;
  const HostTensor *get(const std::string &k) {
    auto it = e->host.find(k);
    return it == e->host.end() ? nullptr : &it->second;
  }
  bool has(const std::string &k) { return get(k) != nullptr; }
  // internal/native/linear.go:19-49 (loadLinear)
  int linear(const std::string &name, bool with_bias, DLinear *l) {
    const HostTensor *w = get(name + ".weight");
    if (!w)
      PTTS_FAIL(PTTS_EINVAL, "native: tensor %s.weight not found", name.c_str());
    if (w->shape.size() != 2)
      PTTS_FAIL(PTTS_EINVAL, "native: linear \"%s\" weight must be rank-2", name.c_str());
    l->out = (int)w->shape[0];
    l->in = (int)w->shape[1];
    int rc = e->upload(w->data, &l->w);
    if (rc)
      return rc;
    if ((rc = e->make_tc(l->w, w->data, l->out, l->in)))
      return rc;
    l->b = nullptr;
    if (with_bias) {
      const HostTensor *b = get(name + ".bias");
      if (b) {
        if (b->shape.size() != 1 || b->shape[0] != w->shape[0])
          PTTS_FAIL(PTTS_EINVAL, "native: linear \"%s\" bias shape incompatible with weight",
                    name.c_str());
        rc = e->upload(b->data, &l->b);
        if (rc)
          return rc;
      }
    }
    return PTTS_OK;
  }
  // internal/native/linear.go:191-208 (loadLayerNorm)
  int lnorm(const std::string &name, float eps, DLayerNorm *ln) {
    const HostTensor *w = get(name + ".weight"), *b = get(name + ".bias");
    if (!w || !b)
      PTTS_FAIL(PTTS_EINVAL, "native: tensor %s.weight/bias not found", name.c_str());
    if (w->shape.size() != 1 || b->shape.size() != 1 || w->shape[0] != b->shape[0])
      PTTS_FAIL(PTTS_EINVAL, "native: layernorm \"%s\" invalid shapes", name.c_str());
    ln->d = (int)w->shape[0];
    ln->eps = eps;
    int rc = e->upload(w->data, &ln->w);
    if (rc)
      return rc;
    return e->upload(b->data, &ln->b);
  }
  int tlayer(const std::string &p, bool mimi, DTLayer *l) {
    int rc;
    if ((rc = lnorm(p + ".norm1", 1e-5f, &l->n1))) return rc;
    if ((rc = lnorm(p + ".norm2", 1e-5f, &l->n2))) return rc;
    if ((rc = linear(p + ".self_attn.in_proj", false, &l->in_proj))) return rc;
    if ((rc = linear(p + ".self_attn.out_proj", false, &l->out_proj))) return rc;
    if ((rc = linear(p + ".linear1", false, &l->l1))) return rc;
    if ((rc = linear(p + ".linear2", false, &l->l2))) return rc;
    if (mimi) {
      if (const HostTensor *t = get(p + ".layer_scale_1.scale"))
        if ((rc = e->upload(t->data, &l->ls1))) return rc;
      if (const HostTensor *t = get(p + ".layer_scale_2.scale"))
        if ((rc = e->upload(t->data, &l->ls2))) return rc;
    }
    return PTTS_OK;
  }
  // Conv1d [Cout,Cin,k] -> GEMM weight [Cout, k*Cin] (tap-major K so that the
  // A row of output t is the contiguous channels-last span t-k+1..t).
  int conv(const std::string &p, DConvG *c) {
    const HostTensor *w = get(p + ".weight");
    if (!w)
      PTTS_FAIL(PTTS_EINVAL, "native: tensor %s.weight not found", p.c_str());
    if (w->shape.size() != 3)
      PTTS_FAIL(PTTS_EINVAL, "native: conv1d weight must be rank-3, got rank %zu",
                w->shape.size());
    int co = (int)w->shape[0], ci = (int)w->shape[1], k = (int)w->shape[2];
    std::vector<float> wp((size_t)co * ci * k);
    for (int oc = 0; oc < co; oc++)
      for (int ic = 0; ic < ci; ic++)
        for (int kx = 0; kx < k; kx++)
          wp[((size_t)oc * k + kx) * ci + ic] = w->data[((size_t)oc * ci + ic) * k + kx];
    c->cin = ci; c->cout = co; c->taps = k; c->stride = 1; c->N = co; c->K = k * ci;
    int rc = e->upload(wp, &c->w);
    if (rc)
      return rc;
    if ((rc = e->make_tc(c->w, wp, c->N, c->K)))
      return rc;
    c->b = nullptr;
    if (const HostTensor *b = get(p + ".bias"))
      if ((rc = e->upload(b->data, &c->b)))
        return rc;
    return PTTS_OK;
  }
  // ConvTranspose1d [Cin,Cout,k], stride s, right trim k-s -> s phase GEMMs
  // stacked along N: W[(r*Cout+oc), (j*Cin+ic)] = w[ic][oc][r + s*(J-1-j)].
  int convtr(const std::string &p, int stride, DConvG *c) {
    const HostTensor *w = get(p + ".weight");
    if (!w)
      PTTS_FAIL(PTTS_EINVAL, "native: tensor %s.weight not found", p.c_str());
    if (w->shape.size() != 3)
      PTTS_FAIL(PTTS_EINVAL, "native: convtranspose1d weight must be rank-3");
    int ci = (int)w->shape[0], co = (int)w->shape[1], k = (int)w->shape[2];
    if (k < stride)
      PTTS_FAIL(PTTS_EINVAL, "native-cuda: convtranspose1d kernel %d < stride %d", k, stride);
    int J = (k + stride - 1) / stride;
    std::vector<float> wp((size_t)stride * co * J * ci, 0.0f);
    for (int r = 0; r < stride; r++)
      for (int oc = 0; oc < co; oc++)
        for (int j = 0; j < J; j++) {
          int kx = r + stride * (J - 1 - j);
          if (kx >= k)
            continue;
          for (int ic = 0; ic < ci; ic++)
            wp[((size_t)(r * co + oc) * J + j) * ci + ic] = w->data[((size_t)ic * co + oc) * k + kx];
        }
    c->cin = ci; c->cout = co; c->taps = J; c->stride = stride; c->N = stride * co; c->K = J * ci;
    int rc = e->upload(wp, &c->w);
    if (rc)
      return rc;
    if ((rc = e->make_tc(c->w, wp, c->N, c->K)))
      return rc;
    c->b = nullptr;
    if (const HostTensor *b = get(p + ".bias")) {
      std::vector<float> bp((size_t)stride * co);
      for (int r = 0; r < stride; r++)
        for (int oc = 0; oc < co; oc++)
          bp[(size_t)r * co + oc] = b->data[oc];
      if ((rc = e->upload(bp, &c->b)))
        return rc;
    }
    return PTTS_OK;
  }
};

// internal/native/flow_transformer.go:797-832 (buildRoPE)
void build_rope(int max_seq, int hd, double max_period, std::vector<float> &c,
                std::vector<float> &s) {
  int half = hd / 2;
  c.resize((size_t)max_seq * half);
  s.resize((size_t)max_seq * half);
  for (int pos = 0; pos < max_seq; pos++)
    for (int i = 0; i < half; i++) {
      double f = 1.0 / std::pow(max_period, (double)i / (double)half);
      double a = (double)pos * f;
      c[(size_t)pos * half + i] = (float)std::cos(a);
      s[(size_t)pos * half + i] = (float)std::sin(a);
    }
}
} // namespace

int Engine::finalize() {
  if (finalized)
    PTTS_FAIL(PTTS_ESTATE, "native-cuda: weights already finalized");
  CUDA_TRY(cudaSetDevice(device));
  Loader ld{this};
  int rc;
  // ---- FlowLM (flow_lm.go:51-119) ----
  const HostTensor *emb = ld.get("flow_lm.conditioner.embed.weight");
  if (!emb || emb->shape.size() != 2)
    PTTS_FAIL(PTTS_EINVAL, "native: load conditioner: embed must be rank-2");
  nbins = (int)emb->shape[0];
  D = (int)emb->shape[1];
  if ((rc = upload(emb->data, &embed))) return rc;
  for (int i = 0;; i++) {
    std::string p = "flow_lm.transformer.layers." + std::to_string(i);
    if (!ld.has(p + ".norm1.weight"))
      break;
    DTLayer l;
    if ((rc = ld.tlayer(p, false, &l))) return rc;
    fl.push_back(l);
  }
  L = (int)fl.size();
  if (L == 0)
    PTTS_FAIL(PTTS_EINVAL, "native: no flow_lm transformer layers found");
  if (fl[0].out_proj.out != D || D % H)
    PTTS_FAIL(PTTS_EINVAL, "native: d_model %d not divisible by num_heads %d", D, H);
  Dh = D / H;
  if (Dh != 64)
    PTTS_FAIL(PTTS_EINVAL, "native-cuda: head_dim %d unsupported (kernels are specialised for 64)", Dh);
  F = fl[0].l1.out;
  for (auto &l : fl)
    if (l.in_proj.out != 3 * D || l.in_proj.in != D || l.out_proj.in != D ||
        l.l1.in != D || l.l2.in != l.l1.out || l.l2.out != D || l.n1.d != D || l.n2.d != D)
      PTTS_FAIL(PTTS_EINVAL, "native-cuda: flow_lm transformer layer shape mismatch");
  {
    const HostTensor *b = ld.get("flow_lm.bos_emb");
    if (!b || b->shape.size() != 1 || b->shape[0] != ldim)
      PTTS_FAIL(PTTS_EINVAL, "native: tensor flow_lm.bos_emb missing or not [%d]", ldim);
    if ((rc = upload(b->data, &bos))) return rc;
  }
  if ((rc = ld.linear("flow_lm.input_linear", true, &input_linear))) return rc;
  if ((rc = ld.lnorm("flow_lm.out_norm", 1e-5f, &out_norm))) return rc;
  if ((rc = ld.linear("flow_lm.out_eos", true, &out_eos_lin))) return rc;
  if (input_linear.in != ldim || input_linear.out != D || out_eos_lin.in != D || out_eos_lin.out != 1)
    PTTS_FAIL(PTTS_EINVAL, "native-cuda: flow_lm input_linear/out_eos shape mismatch");
  if (const HostTensor *b = ld.get("flow_lm.out_eos.bias"))
    out_eos_bias = b->data[0];
  // ---- flow net (flow_net.go:250-305) ----
  if ((rc = ld.linear("flow_lm.flow_net.cond_embed", true, &cond_embed))) return rc;
  if ((rc = ld.linear("flow_lm.flow_net.input_proj", true, &input_proj))) return rc;
  C = input_proj.out;
  std::vector<float> ada_w, ada_b;
  bool ada_has_bias = true;
  auto append_ada = [&](const std::string &name, int want_out) -> int {
    const HostTensor *w = ld.get(name + ".weight"), *b = ld.get(name + ".bias");
    if (!w || w->shape.size() != 2 || w->shape[0] != want_out || w->shape[1] != C)
      PTTS_FAIL(PTTS_EINVAL, "native: flow modulation %s shape invalid", name.c_str());
    ada_w.insert(ada_w.end(), w->data.begin(), w->data.end());
    if (b)
      ada_b.insert(ada_b.end(), b->data.begin(), b->data.end());
    else {
      ada_b.insert(ada_b.end(), (size_t)want_out, 0.0f);
      ada_has_bias = ada_has_bias && false;
    }
    return PTTS_OK;
  };
  for (int i = 0;; i++) {
    std::string p = "flow_lm.flow_net.res_blocks." + std::to_string(i);
    if (!ld.has(p + ".in_ln.weight"))
      break;
    DResBlk b;
    if ((rc = ld.lnorm(p + ".in_ln", 1e-6f, &b.in_ln))) return rc;
    if ((rc = ld.linear(p + ".mlp.0", true, &b.mlp0))) return rc;
    if ((rc = ld.linear(p + ".mlp.2", true, &b.mlp2))) return rc;
    if ((rc = append_ada(p + ".adaLN_modulation.1", 3 * C))) return rc;
    if (b.mlp0.in != C || b.mlp2.out != C || b.mlp2.in != b.mlp0.out || b.in_ln.d != C)
      PTTS_FAIL(PTTS_EINVAL, "native-cuda: flow res block %d shape mismatch", i);
    rb.push_back(b);
  }
  depth = (int)rb.size();
  if (depth == 0)
    PTTS_FAIL(PTTS_EINVAL, "native: no flow_net res blocks found");
  if ((rc = append_ada("flow_lm.flow_net.final_layer.adaLN_modulation.1", 2 * C))) return rc;
  ada_n = depth * 3 * C + 2 * C;
  ada_all.in = C;
  ada_all.out = ada_n;
  if ((rc = upload(ada_w, &ada_all.w))) return rc;
  if ((rc = make_tc(ada_all.w, ada_w, ada_n, C))) return rc;
  if ((rc = upload(ada_b, &ada_all.b))) return rc;
  if ((rc = ld.linear("flow_lm.flow_net.final_layer.linear", true, &final_linear))) return rc;
  if (final_linear.in != C || final_linear.out != ldim || cond_embed.in != D ||
      cond_embed.out != C || input_proj.in != ldim)
    PTTS_FAIL(PTTS_EINVAL, "native-cuda: flow_net shape mismatch");
  for (int i = 0; i < 2; i++) {
    std::string p = "flow_lm.flow_net.time_embed." + std::to_string(i);
    const HostTensor *fq = ld.get(p + ".freqs"), *w1 = ld.get(p + ".mlp.0.weight"),
                     *b1 = ld.get(p + ".mlp.0.bias"), *w2 = ld.get(p + ".mlp.2.weight"),
                     *b2 = ld.get(p + ".mlp.2.bias"), *al = ld.get(p + ".mlp.3.alpha");
    if (!fq || !w1 || !w2 || !al)
      PTTS_FAIL(PTTS_EINVAL, "native: tensor %s.* not found", p.c_str());
    TEmb &t = te[i];
    t.freqs = fq->data;
    t.w1 = w1->data;
    t.w2 = w2->data;
    t.alpha = al->data;
    t.hid = (int)w1->shape[0];
    t.in = (int)w1->shape[1];
    t.out = (int)w2->shape[0];
    t.b1 = b1 ? b1->data : std::vector<float>((size_t)t.hid, 0.0f);
    t.b2 = b2 ? b2->data : std::vector<float>((size_t)t.out, 0.0f);
    if (t.in != 2 * (int)t.freqs.size() || t.out != C || (int)w2->shape[1] != t.hid ||
        (int)t.alpha.size() != C)
      PTTS_FAIL(PTTS_EINVAL, "native-cuda: time_embed.%d shape mismatch", i);
  }
  nfreq = (int)te[0].freqs.size();
  {
    std::vector<float> c, s;
    build_rope(8192, Dh, 10000.0, c, s);
    if ((rc = upload(c, &rope_cos))) return rc;
    if ((rc = upload(s, &rope_sin))) return rc;
  }
  // ---- Mimi (mimi.go:546-637) ----
  {
    const HostTensor *q = ld.get("mimi.quantizer.output_proj.weight");
    const HostTensor *sd = ld.get("flow_lm.emb_std"), *mn = ld.get("flow_lm.emb_mean");
    if (!q || q->shape.size() != 3)
      PTTS_FAIL(PTTS_EINVAL, "native: quantizer projection weight rank must be 3");
    if (!sd || !mn || (int)sd->data.size() < ldim || (int)mn->data.size() < ldim)
      PTTS_FAIL(PTTS_EINVAL, "native: flow emb stats shape mismatch");
    int oc_n = (int)q->shape[0], ic_n = (int)q->shape[1], k = (int)q->shape[2];
    if (k != 1 || ic_n != ldim)
      PTTS_FAIL(PTTS_EINVAL, "native-cuda: quantizer.output_proj must be [Md,%d,1]", ldim);
    // model.go:226-242: W*std, bias = sum W*mean
    std::vector<float> pw((size_t)oc_n * ic_n), pb((size_t)oc_n);
    for (int oc = 0; oc < oc_n; oc++) {
      float bv = 0.0f;
      for (int ic = 0; ic < ic_n; ic++) {
        float w = q->data[(size_t)oc * ic_n + ic];
        pw[(size_t)oc * ic_n + ic] = w * sd->data[ic];
        bv += w * mn->data[ic];
      }
      pb[oc] = bv;
    }
    Md = oc_n;
    proj.in = ic_n;
    proj.out = oc_n;
    if ((rc = upload(pw, &proj.w))) return rc;
    if ((rc = make_tc(proj.w, pw, proj.out, proj.in))) return rc;
    if ((rc = upload(pb, &proj.b))) return rc;
  }
  {
    const HostTensor *u = ld.get("mimi.upsample.convtr.convtr.weight");
    if (!u || u->shape.size() != 3 || u->shape[0] != Md || u->shape[1] != 1)
      PTTS_FAIL(PTTS_EINVAL, "native-cuda: mimi.upsample.convtr must be depthwise [Md,1,k]");
    up_k = (int)u->shape[2];
    up_stride = 16; // mimi.go:25-34,665-686: encoder_frame_rate / frame_rate
    if (up_k < up_stride)
      PTTS_FAIL(PTTS_EINVAL, "native-cuda: upsample kernel %d < stride %d", up_k, up_stride);
    if ((rc = upload(u->data, &up_w))) return rc;
  }
  for (int i = 0;; i++) {
    std::string p = "mimi.decoder_transformer.transformer.layers." + std::to_string(i);
    if (!ld.has(p + ".norm1.weight"))
      break;
    DTLayer l;
    if ((rc = ld.tlayer(p, true, &l))) return rc;
    ml.push_back(l);
  }
  ML = (int)ml.size();
  if (ML == 0)
    PTTS_FAIL(PTTS_EINVAL, "native: no mimi decoder transformer layers found");
  if (ml[0].out_proj.out != Md || Md % MH || Md / MH != 64)
    PTTS_FAIL(PTTS_EINVAL, "native-cuda: mimi d_model %d / heads %d must give head_dim 64", Md, MH);
  MF = ml[0].l1.out;
  {
    std::vector<float> c, s;
    build_rope(MIMI_ROPE_ROWS, Md / MH, 10000.0, c, s);
    if ((rc = upload(c, &mrope_cos))) return rc;
    if ((rc = upload(s, &mrope_sin))) return rc;
  }
  if ((rc = ld.conv("mimi.decoder.model.0.conv", &init_conv))) return rc;
  const int strides[3] = {6, 5, 4};
  const char *upk[3] = {"2", "5", "8"};
  const char *rsk[3] = {"3", "6", "9"};
  int ch = init_conv.cout;
  if (init_conv.cin != Md)
    PTTS_FAIL(PTTS_EINVAL, "native-cuda: seanet init conv expects %d input channels", Md);
  for (int s = 0; s < 3; s++) {
    std::string base = "mimi.decoder.model.";
    if ((rc = ld.convtr(base + upk[s] + ".convtr", strides[s], &up[s]))) return rc;
    if (up[s].cin != ch)
      PTTS_FAIL(PTTS_EINVAL, "native-cuda: seanet stage %d channel mismatch", s);
    ch = up[s].cout;
    if ((rc = ld.conv(base + rsk[s] + ".block.1.conv", &res1[s]))) return rc;
    if ((rc = ld.conv(base + rsk[s] + ".block.3.conv", &res2[s]))) return rc;
    if (res1[s].cin != ch || res2[s].cout != ch || res2[s].cin != res1[s].cout)
      PTTS_FAIL(PTTS_EINVAL, "native-cuda: seanet res block %d channel mismatch", s);
  }
  {
    const HostTensor *w = ld.get("mimi.decoder.model.11.conv.weight");
    const HostTensor *b = ld.get("mimi.decoder.model.11.conv.bias");
    if (!w || w->shape.size() != 3 || w->shape[0] != 1 || w->shape[1] != ch)
      PTTS_FAIL(PTTS_EINVAL, "native-cuda: final conv must be [1,%d,k]", ch);
    final_cin = ch;
    final_k = (int)w->shape[2];
    std::vector<float> wp((size_t)final_cin * final_k);
    for (int ic = 0; ic < final_cin; ic++)
      for (int kx = 0; kx < final_k; kx++)
        wp[(size_t)kx * final_cin + ic] = w->data[(size_t)ic * final_k + kx];
    if ((rc = upload(wp, &final_w))) return rc;
    final_b = b ? b->data[0] : 0.0f;
  }
  if (init_conv.taps - 1 > mimi_pad || res1[0].taps - 1 > mimi_pad)
    PTTS_FAIL(PTTS_EINVAL, "native-cuda: conv kernel wider than the %d-row segment pad", mimi_pad);

  // ---- model info (SURVEY §8d formulas evaluated on the loaded shapes) ----
  info.d_model = D; info.n_heads = H; info.head_dim = Dh; info.n_layers = L; info.ffn_dim = F;
  info.flow_dim = C; info.flow_depth = depth; info.latent_dim = ldim; info.n_bins = nbins;
  info.mimi_dim = Md; info.mimi_heads = MH; info.mimi_layers = ML; info.mimi_ffn_dim = MF;
  info.mimi_context = Mctx; info.samples_per_frame = up_stride * 6 * 5 * 4;
  info.mimi_steps_per_latent = up_stride;
  {
    int64_t tl = (int64_t)L * (4LL * D * D + 2LL * D * F + 4LL * D);
    int64_t io = (int64_t)ldim * D + (input_linear.b ? D : 0) + 2LL * D + D + 1;
    int64_t temb = 0;
    for (int i = 0; i < 2; i++)
      temb += (int64_t)te[i].freqs.size() + (int64_t)te[i].hid * te[i].in + te[i].hid +
              (int64_t)te[i].out * te[i].hid + te[i].out + C;
    int64_t flow = (int64_t)ldim * C + C + (int64_t)D * C + C + temb +
                   (int64_t)depth * (2LL * C + 2LL * ((int64_t)C * C + C) + 3LL * C * C + 3LL * C) +
                   ((int64_t)ldim * C + ldim) + (2LL * C * C + 2LL * C);
    info.params_ar = tl + io + flow;
    info.macs_ar_frame = (int64_t)L * (4LL * D * D + 2LL * D * F) + (int64_t)ldim * D + D +
                         (int64_t)ldim * C + (int64_t)D * C + (int64_t)ada_n * C +
                         (int64_t)depth * 2LL * C * C + (int64_t)ldim * C;
    info.macs_prefill_token = (int64_t)L * (4LL * D * D + 2LL * D * F);
    int64_t mt = (int64_t)ML * (4LL * Md * Md + 2LL * Md * MF);
    int64_t per16 = mt + (int64_t)init_conv.N * init_conv.K + (int64_t)up[0].N * up[0].K;
    int64_t r1 = (int64_t)res1[0].N * res1[0].K + (int64_t)res2[0].N * res2[0].K +
                 (int64_t)up[1].N * up[1].K;
    int64_t r2 = (int64_t)res1[1].N * res1[1].K + (int64_t)res2[1].N * res2[1].K +
                 (int64_t)up[2].N * up[2].K;
    int64_t r3 = (int64_t)res1[2].N * res1[2].K + (int64_t)res2[2].N * res2[2].K +
                 (int64_t)final_cin * final_k;
    info.macs_mimi_frame = (int64_t)ldim * Md + (int64_t)up_stride * (per16 + 6 * (r1 + 5 * (r2 + 4 * r3)));
    int64_t pm = (int64_t)ldim * Md + (int64_t)Md * up_k + (int64_t)ML * (4LL * Md * Md + 2LL * Md * MF + 6LL * Md);
    pm += (int64_t)init_conv.N * init_conv.K + init_conv.N;
    for (int s = 0; s < 3; s++)
      pm += (int64_t)up[s].N * up[s].K / 1 + (int64_t)res1[s].N * res1[s].K + (int64_t)res2[s].N * res2[s].K;
    info.params_mimi = pm;
  }

  // ---- KV pool, slots ----
  max_batch = cfg.max_batch;
  max_frames = cfg.max_frames;
  page_tokens = cfg.page_tokens;
  max_pages = (cfg.max_seq_tokens + page_tokens - 1) / page_tokens;
  if (max_pages > KV_MAX_SEQ_PAGES)
    PTTS_FAIL(PTTS_EINVAL, "native-cuda: max_seq_tokens %d / page_tokens %d = %d pages per sequence (max %d)",
              cfg.max_seq_tokens, page_tokens, max_pages, KV_MAX_SEQ_PAGES);
  n_pages = cfg.kv_pages > 0 ? cfg.kv_pages : max_batch * max_pages;
  // The cache holds K (rotated) and V as fp16 by default: its stream is what bounds the batched AR
  // step (48 KB per token and sequence as fp32), and 11 mantissa bits keep the attention output
  // ~1e-4 relative — inside the latent tolerance; PTTS_ENGINE_KV_F32 keeps the reference's fp32.
  kv_half = !(cfg.flags & PTTS_ENGINE_KV_F32) && !getenv("PTTS_KV_F32");
  size_t page_floats = (size_t)L * 2 * H * page_tokens * Dh;
  if ((rc = dmalloc((void **)&kv_pool, (size_t)n_pages * page_floats * (kv_half ? 2 : 4)))) return rc;
  for (int i = n_pages - 1; i >= 0; i--)
    free_pages.push_back(i);
  for (int i = max_batch - 1; i >= 0; i--)
    free_slots.push_back(i);
  h_pagetab.assign((size_t)max_batch * max_pages, 0);
  if ((rc = dmalloc((void **)&d_pagetab, h_pagetab.size() * sizeof(int32_t)))) return rc;
  CUDA_TRY(cudaMemset(d_pagetab, 0, h_pagetab.size() * sizeof(int32_t)));
  kvg.pool = kv_pool; kvg.half = kv_half ? 1 : 0; kvg.pagetab = d_pagetab; kvg.max_pages = max_pages;
  kvg.page_tokens = page_tokens; kvg.n_layers = L; kvg.n_heads = H; kvg.head_dim = Dh;
  if ((rc = dmalloc((void **)&d_hist, (size_t)(max_batch + 1) * (max_frames + 1) * ldim * sizeof(float)))) return rc;
  if ((rc = dmalloc((void **)&d_noise, (size_t)(max_batch + 1) * max_frames * ldim * sizeof(float)))) return rc;
  CUDA_TRY(cudaMemset(d_noise, 0, (size_t)(max_batch + 1) * max_frames * ldim * sizeof(float)));
  {
    std::vector<float> nanrow((size_t)ldim, NAN);
    CUDA_TRY(cudaMemset(d_hist, 0, (size_t)(max_batch + 1) * (max_frames + 1) * ldim * sizeof(float)));
    for (int s = 0; s <= max_batch; s++)
      CUDA_TRY(cudaMemcpy(d_hist + (size_t)s * (max_frames + 1) * ldim, nanrow.data(),
                          ldim * sizeof(float), cudaMemcpyHostToDevice));
  }
  gemm_ws_floats = 16L << 20;
  if ((rc = dmalloc((void **)&gemm_ws, (size_t)gemm_ws_floats * 4))) return rc;
  if ((rc = dmalloc((void **)&gemm_counters, (size_t)gemm_n_counters * 4))) return rc;
  CUDA_TRY(cudaMemset(gemm_counters, 0, (size_t)gemm_n_counters * 4));
  if ((rc = dmalloc((void **)&gemm_ws_m, (size_t)gemm_ws_floats * 4))) return rc;
  if ((rc = dmalloc((void **)&gemm_counters_m, (size_t)gemm_n_counters * 4))) return rc;
  CUDA_TRY(cudaMemset(gemm_counters_m, 0, (size_t)gemm_n_counters * 4));
  // ---- FlowLM workspaces ----
  max_rows = cfg.max_prefill_tokens > max_batch ? cfg.max_prefill_tokens : max_batch;
  if ((rc = dmalloc((void **)&x, (size_t)max_rows * D * 4))) return rc;
  if ((rc = dmalloc((void **)&nrm, (size_t)max_rows * D * 4))) return rc;
  if ((rc = dmalloc((void **)&att, (size_t)max_rows * D * 4))) return rc;
  if ((rc = dmalloc((void **)&qkv, (size_t)max_rows * 3 * D * 4))) return rc;
  if ((rc = dmalloc((void **)&ff, (size_t)max_rows * F * 4))) return rc;
  if ((rc = dmalloc((void **)&d_row_slot, (size_t)max_rows * 4))) return rc;
  if ((rc = dmalloc((void **)&d_row_pos, (size_t)max_rows * 4))) return rc;
  if ((rc = dmalloc((void **)&d_ids, (size_t)max_rows * 8))) return rc;
  if ((rc = dmalloc((void **)&d_off_a, (size_t)max_rows * 8))) return rc;
  if ((rc = dmalloc((void **)&d_off_b, (size_t)max_rows * 8))) return rc;
  if ((rc = dmalloc((void **)&d_off_c, (size_t)max_rows * 8))) return rc;
  CUDA_TRY(cudaMallocHost((void **)&h_row_slot, (size_t)max_rows * 4));
  CUDA_TRY(cudaMallocHost((void **)&h_row_pos, (size_t)max_rows * 4));
  CUDA_TRY(cudaMallocHost((void **)&h_ids, (size_t)max_rows * 8));
  CUDA_TRY(cudaMallocHost((void **)&h_off_a, (size_t)max_rows * 8));
  CUDA_TRY(cudaMallocHost((void **)&h_off_b, (size_t)max_rows * 8));
  CUDA_TRY(cudaMallocHost((void **)&h_off_c, (size_t)max_rows * 8));
  CUDA_TRY(cudaMallocHost((void **)&h_eos, (size_t)max_batch * 4));
  CUDA_TRY(cudaMallocHost((void **)&h_lat, (size_t)max_batch * ldim * 4));
  CUDA_TRY(cudaMallocHost((void **)&h_stage, (size_t)max_batch * ldim * 4 * 2));
  const int B = max_batch;
  if ((rc = dmalloc((void **)&fh_in, (size_t)B * ldim * 4))) return rc;
  if ((rc = dmalloc((void **)&fh_frames, (size_t)B * ldim * 4))) return rc;
  if ((rc = dmalloc((void **)&fh_noise, (size_t)B * ldim * 4))) return rc;
  if ((rc = dmalloc((void **)&fh_cur, (size_t)B * ldim * 4))) return rc;
  if ((rc = dmalloc((void **)&fh_last, (size_t)B * D * 4))) return rc;
  if ((rc = dmalloc((void **)&fh_eos, (size_t)B * 4))) return rc;
  if ((rc = dmalloc((void **)&fh_cproj, (size_t)B * C * 4))) return rc;
  if ((rc = dmalloc((void **)&fh_sy, (size_t)B * C * 4))) return rc;
  if ((rc = dmalloc((void **)&fh_ada, (size_t)B * ada_n * 4))) return rc;
  if ((rc = dmalloc((void **)&fh_x, (size_t)B * C * 4))) return rc;
  if ((rc = dmalloc((void **)&fh_h, (size_t)B * C * 4))) return rc;
  if ((rc = dmalloc((void **)&fh_h2, (size_t)B * rb[0].mlp0.out * 4))) return rc;
  if ((rc = make_act(s_nrm, max_rows, D, 0))) return rc;
  if ((rc = make_act(s_att, max_rows, D, 0))) return rc;
  if ((rc = make_act(s_ff, max_rows, F, 0))) return rc;
  if ((rc = make_act(s_in, B, ldim, 0))) return rc;
  if ((rc = make_act(s_last, B, D, 0))) return rc;
  if ((rc = make_act(s_sy, B, C, 0))) return rc;
  if ((rc = make_act(s_cur, B, ldim, 0))) return rc;
  if ((rc = make_act(s_h, B, C, 0))) return rc;
  if ((rc = make_act(s_h2, B, rb[0].mlp0.out, 0))) return rc;

  // ---- Mimi workspaces (front slack rows stay zero: causal left context) ----
  mimi_frames_cap = cfg.mimi_group_frames;
  if (mimi_frames_cap < max_frames)
    mimi_frames_cap = max_frames; // one finished chunk must always fit a decode group
  int max_segs = 2 * max_batch + 64;
  if (max_segs > mimi_frames_cap) max_segs = mimi_frames_cap;
  mimi_rows_cap = up_stride * mimi_frames_cap + mimi_pad * max_segs;
  auto slack_alloc = [&](float **p, size_t rows, int chn) -> int {
    float *base = nullptr;
    size_t n = (rows + 8) * (size_t)chn;
    int r = dmalloc((void **)&base, n * sizeof(float));
    if (r)
      return r;
    CUDA_TRY(cudaMemset(base, 0, n * sizeof(float)));
    *p = base + 8 * (size_t)chn;
    return PTTS_OK;
  };
  const size_t Rb = (size_t)mimi_rows_cap;
  if ((rc = dmalloc((void **)&m_latin, (size_t)mimi_frames_cap * ldim * 4))) return rc;
  if ((rc = dmalloc((void **)&m_lat, (size_t)mimi_frames_cap * Md * 4))) return rc;
  if ((rc = slack_alloc(&m_x, Rb, Md))) return rc;
  if ((rc = slack_alloc(&m_nrm, Rb, Md))) return rc;
  if ((rc = slack_alloc(&m_att, Rb, Md))) return rc;
  if ((rc = slack_alloc(&m_qkv, Rb, 3 * Md))) return rc;
  if ((rc = slack_alloc(&m_ff, Rb, MF))) return rc;
  if ((rc = slack_alloc(&m_y0e, Rb, init_conv.cout))) return rc;
  size_t rows = Rb;
  for (int s = 0; s < 3; s++) {
    rows *= (size_t)up[s].stride;
    if ((rc = slack_alloc(&m_y[s], rows, up[s].cout))) return rc;
    if ((rc = slack_alloc(&m_ye[s], rows, up[s].cout))) return rc;
    if ((rc = slack_alloc(&m_he[s], rows, res1[s].cout))) return rc;
  }
  if ((rc = slack_alloc(&m_pcm, rows, 1))) return rc;
  CUDA_TRY(cudaStreamCreateWithFlags(&st_copy, cudaStreamNonBlocking));
  {
    // AR steps are latency-critical chains of short kernels: their stream outranks the Mimi one
    int lo_prio = 0, hi_prio = 0;
    cudaDeviceGetStreamPriorityRange(&lo_prio, &hi_prio);
    CUDA_TRY(cudaStreamCreateWithPriority(&st_mimi, cudaStreamNonBlocking, lo_prio));
    overlap_on = getenv("PTTS_NO_MIMI_OVERLAP") == nullptr;
    mimi_graphs_on = getenv("PTTS_NO_MIMI_GRAPHS") == nullptr;
  }
  for (int i = 0; i < NPCM; i++) {
    if ((rc = slack_alloc(&m_pcm_ring[i], rows, 1))) return rc;
    CUDA_TRY(cudaMallocHost((void **)&h_pcm_ring[i], rows * sizeof(float)));
    if ((rc = dmalloc((void **)&m_s16_ring[i], rows * sizeof(int16_t)))) return rc;
    CUDA_TRY(cudaMallocHost((void **)&h_s16_ring[i], rows * sizeof(int16_t)));
    CUDA_TRY(cudaMallocHost((void **)&h_lat_ring[i], (size_t)mimi_frames_cap * ldim * sizeof(float)));
    if ((rc = dmalloc((void **)&m_latin_ring[i], (size_t)mimi_frames_cap * ldim * 4))) return rc;
    CUDA_TRY(cudaEventCreateWithFlags(&ev_lat_ready[i], cudaEventDisableTiming));
    CUDA_TRY(cudaMallocHost((void **)&h_mrow_pos_ring[i], Rb * 4));
    CUDA_TRY(cudaMallocHost((void **)&h_mrow_frame0_ring[i], Rb * 4));
    CUDA_TRY(cudaMallocHost((void **)&h_mmask_ring[i], Rb));
    CUDA_TRY(cudaMallocHost((void **)&h_moff_ring[i], (size_t)mimi_frames_cap * 8));
    CUDA_TRY(cudaEventCreate(&ev_dec0[i]));
    CUDA_TRY(cudaEventCreate(&ev_pcm_ready[i]));
    CUDA_TRY(cudaEventCreate(&ev_copied[i]));
    CUDA_TRY(cudaEventCreate(&ev_copy0[i]));
  }
  {
    // split planes of every Mimi activation that is a GEMM / conv input (8 zero slack rows
    // in front = the causal left context of the first segment)
    const int lead = 8;
    if ((rc = make_act(sm_latin, mimi_frames_cap, ldim, 0))) return rc;
    if ((rc = make_act(sm_nrm, Rb, Md, lead))) return rc;
    if ((rc = make_act(sm_att, Rb, Md, lead))) return rc;
    if ((rc = make_act(sm_ff, Rb, MF, lead))) return rc;
    if ((rc = make_act(sm_x, Rb, Md, lead))) return rc;
    if ((rc = make_act(sm_y0e, Rb, init_conv.cout, lead))) return rc;
    if ((rc = make_conv_map(sm_x, init_conv.cin, init_conv.taps, &init_conv.amap))) return rc;
    size_t r2 = Rb;
    for (int s = 0; s < 3; s++) {
      const ActSplit &src = s == 0 ? sm_y0e : sm_ye[s - 1];
      if ((rc = make_conv_map(src, up[s].cin, up[s].taps, &up[s].amap))) return rc;
      r2 *= (size_t)up[s].stride;
      if ((rc = make_act(sm_ye[s], r2, up[s].cout, lead))) return rc;
      if ((rc = make_act(sm_he[s], r2, res1[s].cout, lead))) return rc;
      if ((rc = make_conv_map(sm_ye[s], res1[s].cin, res1[s].taps, &res1[s].amap))) return rc;
      if (res2[s].taps != 1)
        if ((rc = make_conv_map(sm_he[s], res2[s].cin, res2[s].taps, &res2[s].amap))) return rc;
      if (s == 1 && up[s].cout == 128) { // the fused block of this stage writes here; up[s+1] reads it
        if ((rc = make_act(sm_yf[s], r2, up[s].cout, lead))) return rc;
        if ((rc = make_conv_map(sm_yf[s], up[s + 1].cin, up[s + 1].taps, &up_alt_amap[s + 1]))) return rc;
      }
    }
  }
  {
    stream_state_on = getenv("PTTS_NO_STREAM_STATE") == nullptr;
    n_stream_states = stream_state_on ? (max_batch < 32 ? max_batch : 32) : 0;
    ring_layer_stride = (long)MIMI_RING * 2 * Md;
    ring_state_stride = (long)ML * ring_layer_stride;
    if (n_stream_states > 0) {
      if ((rc = dmalloc((void **)&mimi_ring, (size_t)n_stream_states * ring_state_stride * sizeof(float)))) return rc;
      CUDA_TRY(cudaMemset(mimi_ring, 0, (size_t)n_stream_states * ring_state_stride * sizeof(float)));
    }
    for (int i = n_stream_states - 1; i >= 0; i--)
      free_stream_states.push_back(i);
    if ((rc = dmalloc((void **)&d_mrow_apos, Rb * 4))) return rc;
    if ((rc = dmalloc((void **)&d_mrow_seg, Rb * sizeof(int4)))) return rc;
    for (int i = 0; i < NPCM; i++) {
      CUDA_TRY(cudaMallocHost((void **)&h_mrow_apos_ring[i], Rb * 4));
      CUDA_TRY(cudaMallocHost((void **)&h_mrow_seg_ring[i], Rb * sizeof(int4)));
    }
  }
  if ((rc = dmalloc((void **)&d_mrow_pos, Rb * 4))) return rc;
  if ((rc = dmalloc((void **)&d_mrow_frame0, Rb * 4))) return rc;
  if ((rc = dmalloc((void **)&d_mmask, Rb))) return rc;
  if ((rc = dmalloc((void **)&d_moff, (size_t)mimi_frames_cap * 8))) return rc;
  CUDA_TRY(cudaMallocHost((void **)&h_mrow_pos, Rb * 4));
  CUDA_TRY(cudaMallocHost((void **)&h_mrow_frame0, Rb * 4));
  CUDA_TRY(cudaMallocHost((void **)&h_mmask, Rb));
  CUDA_TRY(cudaMallocHost((void **)&h_moff, (size_t)mimi_frames_cap * 8));

  if ((rc = setup_fused())) return rc;
  if ((rc = setup_flow_head())) return rc;
  host.clear(); // device copies are authoritative from here on
  finalized = true;
  return PTTS_OK;
}

// ---------------------------------------------------------------------------
// (s,t) schedule: tvec[i] = (TE0(i/steps) + TE1((i+1)/steps)) / 2, computed once
// per lsd_steps on the host exactly as flow_net.go:42-83 + tensor_util.go:273-326.
// ---------------------------------------------------------------------------
static void host_time_embed(const Engine::TEmb &t, float tv, std::vector<float> &out) {
  int nf = (int)t.freqs.size();
  std::vector<float> emb((size_t)2 * nf), h((size_t)t.hid);
  for (int i = 0; i < nf; i++) {
    float arg = tv * t.freqs[i];
    emb[i] = (float)std::cos((double)arg);
    emb[nf + i] = (float)std::sin((double)arg);
  }
  for (int o = 0; o < t.hid; o++) {
    float s = 0.0f;
    for (int i = 0; i < t.in; i++)
      s += emb[i] * t.w1[(size_t)o * t.in + i];
    s += t.b1[o];
    h[o] = s / (1.0f + (float)std::exp((double)(-s)));
  }
  out.assign((size_t)t.out, 0.0f);
  for (int o = 0; o < t.out; o++) {
    float s = 0.0f;
    for (int i = 0; i < t.hid; i++)
      s += h[i] * t.w2[(size_t)o * t.hid + i];
    out[o] = s + t.b2[o];
  }
  double mean = 0;
  for (float v : out)
    mean += (double)v;
  mean /= (double)t.out;
  double var = 0;
  for (float v : out) {
    double d = (double)v - mean;
    var += d * d;
  }
  if (t.out > 1)
    var /= (double)(t.out - 1);
  float inv = (float)(1.0 / std::sqrt(var + (double)1e-5f));
  for (int j = 0; j < t.out; j++)
    out[j] = out[j] * inv * t.alpha[j];
}

int Engine::get_tvec(int lsd_steps, float **out) {
  auto it = tvec_cache.find(lsd_steps);
  if (it != tvec_cache.end()) {
    *out = it->second;
    return PTTS_OK;
  }
  std::vector<float> all((size_t)lsd_steps * C), t0, t1;
  for (int i = 0; i < lsd_steps; i++) {
    float sv = (float)i / (float)lsd_steps, tv = (float)(i + 1) / (float)lsd_steps;
    host_time_embed(te[0], sv, t0);
    host_time_embed(te[1], tv, t1);
    for (int j = 0; j < C; j++)
      all[(size_t)i * C + j] = (t0[j] + t1[j]) * 0.5f;
  }
  float *d = nullptr;
  int rc = upload(all, &d);
  if (rc)
    return rc;
  tvec_cache[lsd_steps] = d;
  *out = d;
  return PTTS_OK;
}

// ---------------------------------------------------------------------------
// sequence resources
// ---------------------------------------------------------------------------
int Engine::seq_alloc(int tokens_needed, Seq **out) {
  if (tokens_needed > cfg.max_seq_tokens)
    PTTS_FAIL(PTTS_ENOMEM, "native-cuda: sequence needs %d KV tokens, max_seq_tokens is %d",
              tokens_needed, cfg.max_seq_tokens);
  int np = (tokens_needed + page_tokens - 1) / page_tokens;
  if (free_slots.empty())
    PTTS_FAIL(PTTS_ENOMEM, "native-cuda: no free sequence slot (max_batch %d)", max_batch);
  if ((int)free_pages.size() < np)
    PTTS_FAIL(PTTS_ENOMEM, "native-cuda: KV pool exhausted (%d pages wanted, %zu free)", np,
              free_pages.size());
  ptts_seq *ps = new ptts_seq();
  Seq *s = &ps->impl;
  s->e = this;
  s->slot = free_slots.back();
  free_slots.pop_back();
  for (int i = 0; i < np; i++) {
    s->pages.push_back(free_pages.back());
    free_pages.pop_back();
  }
  s->kv_cap = np * page_tokens;
  for (int i = 0; i < np; i++)
    h_pagetab[(size_t)s->slot * max_pages + i] = s->pages[i];
  *out = s;
  return sync_pagetab(s);
}

int Engine::sync_pagetab(Seq *s) {
  CUDA_TRY(cudaMemcpyAsync(d_pagetab + (size_t)s->slot * max_pages,
                           h_pagetab.data() + (size_t)s->slot * max_pages,
                           s->pages.size() * sizeof(int32_t), cudaMemcpyHostToDevice, st));
  // h_pagetab is pageable: the copy is staged before return, safe to reuse.
  return PTTS_OK;
}

// ---------------------------------------------------------------------------
// voice prefix cache
// ---------------------------------------------------------------------------
static uint64_t fnv1a(const void *p, size_t n) {
  const unsigned char *c = static_cast<const unsigned char *>(p);
  uint64_t h = 1469598103934665603ull;
  for (size_t i = 0; i < n; i++) {
    h ^= c[i];
    h *= 1099511628211ull;
  }
  return h;
}

// Lookup cost matters: it runs on the scheduler thread for every request.  The key is a
// hash of ~2048 strided samples; a candidate is confirmed by a full compare unless the caller
// passes the very buffer that was confirmed before (the usual server case: one loaded voice
// file serving many requests).
VoiceEntry *Engine::voice_lookup(const float *emb, int frames, bool create) {
  if (!voice_share_on || !emb || frames < page_tokens)
    return nullptr;
  const size_t n = (size_t)frames * D;
  const size_t stride = n / 2048 > 0 ? n / 2048 : 1;
  uint64_t h = 1469598103934665603ull ^ (uint64_t)n;
  for (size_t i = 0; i < n; i += stride) {
    uint32_t u;
    memcpy(&u, emb + i, 4);
    h = (h ^ u) * 1099511628211ull;
  }
  for (auto &v : voice_cache)
    if (v->hash == h && v->frames == frames &&
        (v->src == emb || memcmp(v->data.data(), emb, n * sizeof(float)) == 0)) {
      v->src = emb;
      v->last_use = ++voice_clock;
      return v.get();
    }
  if (!create)
    return nullptr;
  while ((int)voice_cache.size() >= voice_cache_max)
    if (!voice_evict_one())
      return nullptr;
  auto v = std::make_unique<VoiceEntry>();
  v->hash = h;
  v->frames = frames;
  v->src = emb;
  v->data.assign(emb, emb + n);
  v->last_use = ++voice_clock;
  voice_cache.push_back(std::move(v));
  return voice_cache.back().get();
}

// Drops the least recently used entry that no live sequence maps; its pages return to the pool.
bool Engine::voice_evict_one() {
  int best = -1;
  for (size_t i = 0; i < voice_cache.size(); i++)
    if (voice_cache[i]->refs == 0 && (best < 0 || voice_cache[i]->last_use < voice_cache[best]->last_use))
      best = (int)i;
  if (best < 0)
    return false;
  for (int p : voice_cache[best]->pages)
    free_pages.push_back(p);
  voice_cache.erase(voice_cache.begin() + best);
  return true;
}

// A sequence whose first pages are the entry's shared pages (read-only) followed by its own.
int Engine::seq_alloc_shared(int tokens_needed, VoiceEntry *v, Seq **out) {
  if (tokens_needed > cfg.max_seq_tokens)
    PTTS_FAIL(PTTS_ENOMEM, "native-cuda: sequence needs %d KV tokens, max_seq_tokens is %d", tokens_needed,
              cfg.max_seq_tokens);
  const int ns = (int)v->pages.size();
  int np = (tokens_needed + page_tokens - 1) / page_tokens - ns;
  if (np < 0)
    np = 0;
  if (free_slots.empty())
    PTTS_FAIL(PTTS_ENOMEM, "native-cuda: no free sequence slot (max_batch %d)", max_batch);
  if ((int)free_pages.size() < np)
    PTTS_FAIL(PTTS_ENOMEM, "native-cuda: KV pool exhausted (%d pages wanted, %zu free)", np, free_pages.size());
  ptts_seq *ps = new ptts_seq();
  Seq *s = &ps->impl;
  s->e = this;
  s->slot = free_slots.back();
  free_slots.pop_back();
  s->voice = v;
  s->n_shared = ns;
  v->refs++;
  s->pages = v->pages;
  for (int i = 0; i < np; i++) {
    s->pages.push_back(free_pages.back());
    free_pages.pop_back();
  }
  s->kv_cap = (int)s->pages.size() * page_tokens;
  s->kv_len = ns * page_tokens; // the shared positions are already in the cache
  for (size_t i = 0; i < s->pages.size(); i++)
    h_pagetab[(size_t)s->slot * max_pages + i] = s->pages[i];
  *out = s;
  voice_hits++;
  voice_shared_tokens += s->kv_len;
  return sync_pagetab(s);
}

// The first sequence of a voice hands its full voice pages to the cache entry once its
// prefill has been issued (later users are ordered after it on the stream).
void Engine::voice_donate(Seq *s, VoiceEntry *v) {
  if (!s || !v || v->ready || s->voice)
    return;
  const int ns = v->frames / page_tokens;
  if (ns <= 0 || (int)s->pages.size() < ns)
    return;
  v->pages.assign(s->pages.begin(), s->pages.begin() + ns);
  v->ready = true;
  v->refs++;
  s->voice = v;
  s->n_shared = ns;
}

void Engine::seq_free(Seq *s) {
  if (!s || s->slot < 0)
    return;
  if (s->voice) {
    s->voice->refs--;
    s->voice = nullptr;
  }
  for (size_t i = (size_t)s->n_shared; i < s->pages.size(); i++)
    free_pages.push_back(s->pages[i]);
  s->n_shared = 0;
  free_slots.push_back(s->slot);
  s->slot = -1;
  s->pages.clear();
  return;
  free_slots.push_back(s->slot);
  s->slot = -1;
  s->pages.clear();
}

// Parameter block + launch geometry of the single-kernel B=1 frame.
int Engine::setup_fused() {
  fused_on = false;
  voice_share_on = getenv("PTTS_NO_VOICE_SHARING") == nullptr;
  if (getenv("PTTS_NO_FUSED_STEP") || !kv_half /* the frame kernel reads the fp16 cache */ || L > FUSED_MAX_LAYERS || depth > FUSED_MAX_LAYERS || Dh != 64 || D % 4 || C % 4 ||
      ldim > 128 || ldim % 4 || F % 4 || rb[0].mlp0.out % 4 || ada_n % 2)
    return PTTS_OK;
  FusedParams hp{};
  hp.D = D; hp.H = H; hp.F = F; hp.L = L; hp.C = C; hp.hid = rb[0].mlp0.out; hp.depth = depth; hp.ldim = ldim;
  hp.ada_n = ada_n;
  for (int l = 0; l < L; l++) {
    const DTLayer &t = fl[l];
    hp.layer[l] = FusedLayer{t.n1.w, t.n1.b, t.in_proj.w, t.out_proj.w, t.n2.w, t.n2.b, t.l1.w, t.l2.w, t.n1.eps, t.n2.eps};
  }
  hp.bos = bos; hp.in_w = input_linear.w; hp.in_b = input_linear.b;
  hp.on_w = out_norm.w; hp.on_b = out_norm.b; hp.on_eps = out_norm.eps;
  hp.eos_w = out_eos_lin.w; hp.eos_b = out_eos_bias;
  hp.ce_w = cond_embed.w; hp.ce_b = cond_embed.b;
  hp.ada_w = ada_all.w; hp.ada_b = ada_all.b;
  hp.ip_w = input_proj.w; hp.ip_b = input_proj.b;
  for (int b = 0; b < depth; b++) {
    if (rb[b].mlp0.out != hp.hid || rb[b].mlp2.in != hp.hid)
      return PTTS_OK;
    hp.rb[b] = FusedRes{rb[b].in_ln.w, rb[b].in_ln.b, rb[b].mlp0.w, rb[b].mlp0.b, rb[b].mlp2.w, rb[b].mlp2.b, rb[b].in_ln.eps};
  }
  hp.fin_w = final_linear.w; hp.fin_b = final_linear.b;
  hp.rope_cos = rope_cos; hp.rope_sin = rope_sin;
  hp.kv = kvg;
  hp.x = x; hp.qkv = qkv; hp.att = att; hp.ff = ff; hp.fh_in = fh_in; hp.fh_last = fh_last; hp.fh_eos = fh_eos;
  hp.fh_cproj = fh_cproj; hp.fh_ada = fh_ada; hp.fh_x = fh_x; hp.fh_h2 = fh_h2; hp.fh_cur = fh_cur;
  hp.fh_frames = fh_frames; hp.fh_noise = fh_noise;
  hp.hist = d_hist; hp.noise_pool = d_noise; hp.hist_w = d_hist;
  int rc;
  if ((rc = dmalloc((void **)&d_fused_bar, STEP_FUSED_BARRIER_BYTES))) return rc;
  if ((rc = dmalloc((void **)&d_fused_err, 256))) return rc;
  CUDA_TRY(cudaMemset(d_fused_bar, 0, STEP_FUSED_BARRIER_BYTES));
  CUDA_TRY(cudaMemset(d_fused_err, 0, 256));
  CUDA_TRY(cudaMallocHost((void **)&h_fused_err, sizeof(int)));
  *h_fused_err = 0;
  hp.barrier = d_fused_bar;
  hp.error = d_fused_err;
  // which frame kernel: PTTS_FUSED_LL=1 flag-in-data exchange, =0 grid barriers
  {
    const char *v = getenv("PTTS_FUSED_LL");
    fused_ll = v ? atoi(v) != 0 : FUSED_LL_DEFAULT;
  }
  {
    const size_t words = step_fused_ll_words(hp);
    uint2 *llbuf = nullptr;
    if ((rc = dmalloc((void **)&llbuf, (words + 2) * sizeof(uint2)))) return rc;
    CUDA_TRY(cudaMemset(llbuf, 0, (words + 2) * sizeof(uint2)));
    step_fused_ll_assign(hp, llbuf);
    hp.epoch = reinterpret_cast<unsigned *>(llbuf + words); // first tag 1: zeroed words never match
    const unsigned one = 1;
    CUDA_TRY(cudaMemcpy(hp.epoch, &one, sizeof one, cudaMemcpyHostToDevice));
    const char *ns = getenv("PTTS_FUSED_POLL_NS");
    hp.poll_ns = ns ? atoi(ns) : 0;
    const char *la = getenv("PTTS_FUSED_L2_AHEAD");
    hp.l2_ahead = la ? atoi(la) : FUSED_L2_AHEAD_DEFAULT;
    if (hp.l2_ahead < 0 || hp.l2_ahead > 8)
      hp.l2_ahead = 0;
  }
  hp.dbg = nullptr;
  if (getenv("PTTS_FUSED_TIMELINE")) {
    if ((rc = dmalloc((void **)&fused_dbg, 512 * sizeof(unsigned long long)))) return rc;
    CUDA_TRY(cudaMemset(fused_dbg, 0, 512 * sizeof(unsigned long long)));
    hp.dbg = fused_dbg;
  }
  int hidmax = hp.hid > F ? hp.hid : F;
  FusedParams sized = hp;
  sized.F = hidmax;
  fused_smem = step_fused_smem(sized, cfg.max_seq_tokens);
  fused_grid = step_fused_grid(fused_smem, fused_ll);
  if (fused_grid <= 0)
    return PTTS_OK; // cannot be resident: keep the multi-kernel path
  {
    // the weight-ring kernel: bulk copies need 16-byte aligned rows
    const char *v = getenv("PTTS_FUSED_RING");
    fused_ring = v ? atoi(v) != 0 : FUSED_RING_DEFAULT;
    auto al = [](const float *p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
    bool ok = al(hp.ce_w) && al(hp.ada_w) && al(hp.fin_w);
    for (int l = 0; l < L; l++)
      ok = ok && al(hp.layer[l].in_proj) && al(hp.layer[l].out_proj) && al(hp.layer[l].l1) && al(hp.layer[l].l2);
    for (int b = 0; b < depth; b++)
      ok = ok && al(hp.rb[b].mlp0w) && al(hp.rb[b].mlp2w);
    if (fused_ring && ok)
      ring_grid = step_ring_grid(fused_smem, &ring_slots, &ring_smem);
    if (!ok || ring_grid <= 0)
      fused_ring = false;
  }
  int coop = 0;
  cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, device);
  if (!coop)
    return PTTS_OK;
  if ((rc = dmalloc((void **)&d_fused, sizeof(FusedParams)))) return rc;
  CUDA_TRY(cudaMemcpy(d_fused, &hp, sizeof(FusedParams), cudaMemcpyHostToDevice));
  fused_on = true;
  return PTTS_OK;
}

// ---------------------------------------------------------------------------
// FlowLM transformer over R rows already in x; row_slot/row_pos uploaded.
// (flow_transformer.go:326-389, one pass of forwardWithState per layer)
// ---------------------------------------------------------------------------
int Engine::flow_forward(int R, int max_keys, bool decode) {
  static const bool no_fuse = getenv("PTTS_NO_FUSED_ROPE_ATTN") != nullptr;
  decode = decode && !no_fuse;
  int64_t &nl = times.kernel_launches;
  double prof_keys = 0; // sum over rows of keys attended (K and V rows read)
  if (prof_on)
    for (int r = 0; r < R; r++)
      prof_keys += h_row_pos[r] + 1;
  for (int l = 0; l < L; l++) {
    DTLayer &ly = fl[l];
    PK("layernorm", 0, 0, layernorm(x, D, nrm, D, R, D, ly.n1.w, ly.n1.b, ly.n1.eps, nullptr, nullptr, 0, st, s_nrm.out()));
    GemmArgs g;
    g.A = nrm; g.lda = D; g.W = ly.in_proj.w; g.ldw = D; g.M = R; g.N = 3 * D; g.K = D;
    g.epi.C = qkv; g.epi.ldc = 3 * D;
    use_act(g, s_nrm);
    nl += pgemm(g);
    // decode step: RoPE + KV append + attention are one kernel; the prefill (several rows of a
    // sequence per launch) appends first, then attends
    if (!decode)
      PK("rope_kv_append", 0, 0, rope_kv_append(qkv, R, d_row_slot, d_row_pos, rope_cos, rope_sin, kvg, l, st));
    auto attend = [&]() {
      return decode ? attention_paged_decode(qkv, 3 * D, att, D, R, d_row_slot, d_row_pos, rope_cos, rope_sin, kvg, l,
                                             max_keys, st, s_att.out())
                    : attention_paged(qkv, 3 * D, att, D, R, d_row_slot, d_row_pos, kvg, l, max_keys, st, s_att.out());
    };
    if (prof_on) {
      cudaEvent_t a, b;
      prof_begin(&a, &b);
      nl += attend();
      prof_end(a, b, "attention_paged", 4.0 * prof_keys * D, (kv_half ? 4.0 : 8.0) * prof_keys * D);
    } else
      nl += attend();
    GemmArgs o;
    o.A = att; o.lda = D; o.W = ly.out_proj.w; o.ldw = D; o.M = R; o.N = D; o.K = D;
    o.epi.residual = x; o.epi.ldr = D; o.epi.C = x; o.epi.ldc = D;
    use_act(o, s_att);
    nl += pgemm(o);
    PK("layernorm", 0, 0, layernorm(x, D, nrm, D, R, D, ly.n2.w, ly.n2.b, ly.n2.eps, nullptr, nullptr, 0, st, s_nrm.out()));
    GemmArgs f1;
    f1.A = nrm; f1.lda = D; f1.W = ly.l1.w; f1.ldw = D; f1.M = R; f1.N = F; f1.K = D;
    f1.epi.act = ACT_GELU; f1.epi.C = ff; f1.epi.ldc = F; f1.epi.Cs = s_ff.out();
    use_act(f1, s_nrm);
    nl += pgemm(f1);
    GemmArgs f2;
    f2.A = ff; f2.lda = F; f2.W = ly.l2.w; f2.ldw = F; f2.M = R; f2.N = D; f2.K = F;
    f2.epi.residual = x; f2.epi.ldr = D; f2.epi.C = x; f2.epi.ldc = D;
    use_act(f2, s_ff);
    nl += pgemm(f2);
  }
  CUDA_TRY(cudaGetLastError());
  return PTTS_OK;
}

// flow_transformer.go:451-480: import upstream voice KV at positions [0, offset)
int Engine::import_voice_kv(Seq *s, const ptts_voice_kv *vk) {
  if (vk->n_layers != L)
    PTTS_FAIL(PTTS_EINVAL, "native: voice model state has %d layers, transformer has %d",
              vk->n_layers, L);
  int off0 = vk->layers[0].offset;
  for (int l = 0; l < L; l++) {
    const ptts_voice_layer &vl = vk->layers[l];
    if (!vl.cache)
      PTTS_FAIL(PTTS_EINVAL, "native: voice model state module \"transformer.layers.%d.self_attn\" missing cache", l);
    if (vl.offset < 0)
      PTTS_FAIL(PTTS_EINVAL, "native: voice model state module has negative offset %d", vl.offset);
    if (vl.offset > vl.cache_tokens)
      PTTS_FAIL(PTTS_EINVAL, "native: voice model state offset %d exceeds cache length %d",
                vl.offset, vl.cache_tokens);
    if (vl.offset != off0)
      PTTS_FAIL(PTTS_EINVAL, "native-cuda: voice model state offsets differ across layers");
  }
  if (off0 > s->kv_cap)
    PTTS_FAIL(PTTS_ENOMEM, "native-cuda: voice state (%d tokens) exceeds reserved KV", off0);
  if (off0 > max_rows)
    PTTS_FAIL(PTTS_ENOMEM, "native-cuda: voice state (%d tokens) exceeds workspace rows", off0);
  for (int l = 0; l < L; l++) {
    const ptts_voice_layer &vl = vk->layers[l];
    // [2,1,T,H,Dh]: K half rows then V half rows, each row = [H*Dh]
    const float *kh = vl.cache, *vh = vl.cache + (size_t)vl.cache_tokens * D;
    CUDA_TRY(cudaMemcpyAsync(nrm, kh, (size_t)off0 * D * 4, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(att, vh, (size_t)off0 * D * 4, cudaMemcpyHostToDevice, st));
    times.kernel_launches += kv_write_rows(nrm, att, off0, s->slot, 0, kvg, l, st);
    CUDA_TRY(cudaStreamSynchronize(st)); // staging buffers are reused per layer
  }
  s->kv_len = off0;
  return PTTS_OK;
}

// Batched ragged prefill (flow_lm.go:155-187 -> flow_transformer.go:749-771).
int Engine::prefill(const std::vector<Seq *> &seqs, const std::vector<const int64_t *> &tok,
                    const std::vector<int> &ntok, const std::vector<const float *> &voice,
                    const std::vector<int> &nvoice) {
  int R = 0, max_keys = 0;
  for (size_t i = 0; i < seqs.size(); i++) {
    int t = ntok[i] + nvoice[i];
    if (seqs[i]->kv_len + t > seqs[i]->kv_cap)
      PTTS_FAIL(PTTS_ENOMEM, "native-cuda: prompt (%d tokens) exceeds reserved KV", t);
    R += t;
  }
  if (R == 0)
    return PTTS_OK;
  if (R > max_rows)
    PTTS_FAIL(PTTS_ENOMEM, "native-cuda: prefill of %d tokens exceeds max_prefill_tokens %d", R, max_rows);
  cudaEventRecord(ev[0], st);
  int r = 0;
  for (size_t i = 0; i < seqs.size(); i++) {
    Seq *s = seqs[i];
    for (int j = 0; j < nvoice[i]; j++, r++) {
      h_ids[r] = -1;
      h_row_slot[r] = s->slot;
      h_row_pos[r] = s->kv_len + j;
    }
    for (int j = 0; j < ntok[i]; j++, r++) {
      int64_t id = tok[i][j];
      if (id < 0 || id >= nbins)
        PTTS_FAIL(PTTS_EINVAL, "generate: text embeddings: native: token id %d (%lld) out of range [0,%d)",
                  j, (long long)id, nbins);
      h_ids[r] = id;
      h_row_slot[r] = s->slot;
      h_row_pos[r] = s->kv_len + nvoice[i] + j;
    }
    int end = s->kv_len + nvoice[i] + ntok[i];
    if (end > max_keys)
      max_keys = end;
  }
  CUDA_TRY(cudaMemcpyAsync(d_ids, h_ids, (size_t)R * 8, cudaMemcpyHostToDevice, st));
  CUDA_TRY(cudaMemcpyAsync(d_row_slot, h_row_slot, (size_t)R * 4, cudaMemcpyHostToDevice, st));
  CUDA_TRY(cudaMemcpyAsync(d_row_pos, h_row_pos, (size_t)R * 4, cudaMemcpyHostToDevice, st));
  times.kernel_launches += embed_gather(d_ids, embed, x, R, D, st);
  r = 0;
  for (size_t i = 0; i < seqs.size(); i++) {
    if (nvoice[i] > 0)
      CUDA_TRY(cudaMemcpyAsync(x + (size_t)r * D, voice[i], (size_t)nvoice[i] * D * 4,
                               cudaMemcpyHostToDevice, st));
    r += nvoice[i] + ntok[i];
  }
  int rc = flow_forward(R, max_keys, false);
  if (rc)
    return rc;
  for (size_t i = 0; i < seqs.size(); i++)
    seqs[i]->kv_len += nvoice[i] + ntok[i];
  cudaEventRecord(ev[1], st);
  CUDA_TRY(cudaStreamSynchronize(st));
  prof_collect();
  float ms = 0;
  cudaEventElapsedTime(&ms, ev[0], ev[1]);
  times.prefill_ms += ms;
  times.prefill_tokens += R;
  return PTTS_OK;
}

// Parameter block of the persistent flow-head kernel: the phase list of ONE LSD step (the
// reference's default, config.go LSDDecodeSteps = 1); more steps take the multi-kernel path.
int Engine::setup_flow_head() {
  fh_on = false;
  // Opt-in (PTTS_FUSED_HEAD=1).  Measured on B200 (profiles/flow_head_r02.txt): at B = 256 the kernel takes
  // ~380 us against ~200 us for the same head as 27 launches — 22 grid barriers (2.4 us each, plus the wait
  // for the slowest CTA of every phase), GEMM phases without the cluster split-K the stand-alone kernel
  // uses for these tiny shapes, and ~40 us of cooperative-launch overhead.  The multi-kernel head stays
  // the default; the kernel is kept, parity-tested, for the follow-up (split-K inside a cluster launch).
  if (!tc_on || !getenv("PTTS_FUSED_HEAD") || D != 1024 || C != 512 || ldim != 32 || ada_n != depth * 3 * C + 2 * C ||
      2 + 3 * depth + 3 > FH_MAX_PHASES)
    return PTTS_OK;
  const int hid = rb[0].mlp0.out;
  for (auto &b : rb)
    if (b.mlp0.out != hid || b.mlp2.in != hid || hid % 64)
      return PTTS_OK;
  auto tw = [&](const float *w) -> const TcWeight * {
    auto it = tc_map.find(w);
    return it == tc_map.end() ? nullptr : it->second;
  };
  int rc;
  // device copies of the TMA descriptors the kernel dereferences
  auto dev_map = [&](const void *blob, const CUtensorMap **out) -> int {
    void *d = nullptr;
    int r = dmalloc(&d, 128);
    if (r)
      return r;
    CUDA_TRY(cudaMemcpy(d, blob, 128, cudaMemcpyHostToDevice));
    *out = static_cast<const CUtensorMap *>(d);
    return PTTS_OK;
  };
  FhParams hp{};
  hp.D = D; hp.C = C; hp.ldim = ldim; hp.ada_n = ada_n;
  hp.debug = getenv("PTTS_FH_DEBUG") ? atoi(getenv("PTTS_FH_DEBUG")) : 0;
  hp.x = x; hp.noise = fh_noise; hp.last = fh_last; hp.eos = fh_eos; hp.cur = fh_cur; hp.xf = fh_x; hp.ada = fh_ada;
  hp.s_last = s_last.out(); hp.s_cur = s_cur.out(); hp.s_h = s_h.out();
  hp.on_w = out_norm.w; hp.on_b = out_norm.b; hp.on_eps = out_norm.eps;
  hp.eos_w = out_eos_lin.w; hp.eos_b = out_eos_bias;
  hp.ip_w = input_proj.w; hp.ip_b = input_proj.b;
  // cond_embed bias + (TE0(0) + TE1(1)) / 2 of the single Euler step
  {
    float *tv = nullptr;
    if ((rc = get_tvec(1, &tv))) return rc;
    std::vector<float> t((size_t)C), cb((size_t)C, 0.0f);
    CUDA_TRY(cudaMemcpy(t.data(), tv, (size_t)C * 4, cudaMemcpyDeviceToHost));
    if (cond_embed.b)
      CUDA_TRY(cudaMemcpy(cb.data(), cond_embed.b, (size_t)C * 4, cudaMemcpyDeviceToHost));
    for (int i = 0; i < C; i++)
      cb[(size_t)i] += t[(size_t)i];
    if ((rc = upload(cb, &fh_cond_bias))) return rc;
  }
  int np = 0;
  bool ok = true;
  auto add_gemm = [&](const float *w, const ActSplit &a, int N, int K, int bn, const GemmEpi &e) -> int {
    const TcWeight *t = tw(w);
    if (!t || !a.map) {
      ok = false;
      return PTTS_OK;
    }
    FhPhase &f = hp.phase[np++];
    f.kind = FH_GEMM;
    const void *blob = bn == 128 ? t->map128 : (bn == 64 ? t->map64 : t->map32);
    int r;
    if ((r = dev_map(blob, &f.g.wmap))) return r;
    if ((r = dev_map(a.map, &f.g.amap))) return r;
    f.g.n_rows_hi = t->N;
    f.g.a_lo_row = (int)a.plane_rows;
    f.g.N = N; f.g.K = K; f.g.bn = bn;
    f.g.epi = e;
    f.g.epi_kind = flow_head_epi_kind(e);
    return PTTS_OK;
  };
  hp.phase[np++].kind = FH_ROWS_HEAD;
  {
    GemmEpi e;
    e.bias = fh_cond_bias; e.act = ACT_SILU; e.Cs = s_sy.out();
    if ((rc = add_gemm(cond_embed.w, s_last, C, D, 32, e))) return rc;
  }
  {
    GemmEpi e;
    e.bias = ada_all.b; e.C = fh_ada; e.ldc = ada_n;
    if ((rc = add_gemm(ada_all.w, s_sy, ada_n, C, 128, e))) return rc;
  }
  for (int b = 0; b < depth && ok; b++) {
    FhPhase &f = hp.phase[np++];
    f.kind = FH_ROWS_LN;
    f.ln_w = rb[b].in_ln.w; f.ln_b = rb[b].in_ln.b; f.ln_eps = rb[b].in_ln.eps;
    f.shift_off = b * 3 * C; f.scale_off = b * 3 * C + C;
    GemmEpi e0;
    e0.bias = rb[b].mlp0.b; e0.act = ACT_SILU; e0.Cs = s_h2.out();
    if ((rc = add_gemm(rb[b].mlp0.w, s_h, hid, C, 32, e0))) return rc;
    GemmEpi e2;
    e2.bias = rb[b].mlp2.b; e2.mul = fh_ada + (size_t)b * 3 * C + 2 * C; e2.ldmul = ada_n;
    e2.residual = fh_x; e2.ldr = C; e2.C = fh_x; e2.ldc = C;
    if ((rc = add_gemm(rb[b].mlp2.w, s_h2, C, hid, 32, e2))) return rc;
  }
  if (ok) {
    FhPhase &f = hp.phase[np++];
    f.kind = FH_ROWS_LN;
    f.ln_w = nullptr; f.ln_b = nullptr; f.ln_eps = 1e-6f;
    f.shift_off = depth * 3 * C; f.scale_off = depth * 3 * C + C;
    GemmEpi e;
    e.bias = final_linear.b; e.residual = fh_cur; e.ldr = ldim; e.C = fh_cur; e.ldc = ldim;
    if ((rc = add_gemm(final_linear.w, s_h, ldim, C, 32, e))) return rc;
  }
  if (!ok)
    return PTTS_OK; // some weight has no tensor-core copy: keep the multi-kernel head
  hp.n_phases = np;
  fh_grid = flow_head_grid();
  int coop = 0;
  cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, device);
  if (fh_grid <= 0 || !coop)
    return PTTS_OK;
  if ((rc = dmalloc((void **)&d_fh_bar, FH_BARRIER_BYTES))) return rc;
  if ((rc = dmalloc((void **)&d_fh_err, 256))) return rc;
  CUDA_TRY(cudaMemset(d_fh_err, 0, 256));
  CUDA_TRY(cudaMallocHost((void **)&h_fh_err, sizeof(int)));
  *h_fh_err = 0;
  hp.barrier = d_fh_bar;
  hp.error = d_fh_err;
  if ((rc = dmalloc((void **)&d_fh, sizeof(FhParams)))) return rc;
  CUDA_TRY(cudaMemcpy(d_fh, &hp, sizeof(FhParams), cudaMemcpyHostToDevice));
  fh_on = true;
  return PTTS_OK;
}

int Engine::flow_head_fused(int B) {
  const int n = flow_head_launch(d_fh, fh_grid, B, d_fh_bar, st);
  if (n < 0) {
    cudaGetLastError();
    PTTS_FAIL(PTTS_ECUDA, "cuda: cooperative launch of the flow-head kernel failed");
  }
  times.kernel_launches += n;
  fh_used = true;
  return PTTS_OK;
}

// flow_lm.go:262-353: out_norm, EOS logit, LSD Euler decode with the flow net.
// Input: x [B, D] (transformer output), fh_noise [B, ldim].  Output fh_cur.
int Engine::flow_head(int B, int lsd_steps) {
  int64_t &nl = times.kernel_launches;
  float *tvec = nullptr;
  int rc = get_tvec(lsd_steps, &tvec);
  if (rc)
    return rc;
  PK("layernorm", 0, 0, layernorm(x, D, fh_last, D, B, D, out_norm.w, out_norm.b, out_norm.eps, nullptr, nullptr, 0, st, s_last.out()));
  PK("rowdot", 0, 0, rowdot(fh_last, D, out_eos_lin.w, out_eos_bias, B, D, fh_eos, st));
  {
    GemmArgs g;
    g.A = fh_last; g.lda = D; g.W = cond_embed.w; g.ldw = D; g.M = B; g.N = C; g.K = D;
    g.epi.bias = cond_embed.b; g.epi.C = fh_cproj; g.epi.ldc = C;
    use_act(g, s_last);
    nl += pgemm(g);
  }
  PK("elementwise", 0, 0, copy_rows(fh_noise, fh_cur, B, ldim, st, s_cur.out()));
  const float inv = 1.0f / (float)lsd_steps;
  const int hid = rb[0].mlp0.out;
  for (int i = 0; i < lsd_steps; i++) {
    PK("elementwise", 0, 0, add_vec_silu(fh_cproj, tvec + (size_t)i * C, fh_sy, B, C, st, s_sy.out()));
    {
      GemmArgs g;
      g.A = fh_sy; g.lda = C; g.W = ada_all.w; g.ldw = C; g.M = B; g.N = ada_n; g.K = C;
      g.epi.bias = ada_all.b; g.epi.C = fh_ada; g.epi.ldc = ada_n;
      use_act(g, s_sy);
      nl += pgemm(g);
    }
    {
      GemmArgs g;
      g.A = fh_cur; g.lda = ldim; g.W = input_proj.w; g.ldw = ldim; g.M = B; g.N = C; g.K = ldim;
      g.epi.bias = input_proj.b; g.epi.C = fh_x; g.epi.ldc = C;
      use_act(g, s_cur);
      nl += pgemm(g);
    }
    for (int b = 0; b < depth; b++) {
      const float *shift = fh_ada + (size_t)b * 3 * C, *scale = shift + C, *gate = shift + 2 * C;
      PK("layernorm", 0, 0, layernorm(fh_x, C, fh_h, C, B, C, rb[b].in_ln.w, rb[b].in_ln.b, rb[b].in_ln.eps,
                      scale, shift, ada_n, st, s_h.out()));
      GemmArgs g0;
      g0.A = fh_h; g0.lda = C; g0.W = rb[b].mlp0.w; g0.ldw = C; g0.M = B; g0.N = hid; g0.K = C;
      g0.epi.bias = rb[b].mlp0.b; g0.epi.act = ACT_SILU; g0.epi.C = fh_h2; g0.epi.ldc = hid;
      g0.epi.Cs = s_h2.out();
      use_act(g0, s_h);
      nl += pgemm(g0);
      GemmArgs g2;
      g2.A = fh_h2; g2.lda = hid; g2.W = rb[b].mlp2.w; g2.ldw = hid; g2.M = B; g2.N = C; g2.K = hid;
      g2.epi.bias = rb[b].mlp2.b; g2.epi.mul = gate; g2.epi.ldmul = ada_n;
      g2.epi.residual = fh_x; g2.epi.ldr = C; g2.epi.C = fh_x; g2.epi.ldc = C;
      use_act(g2, s_h2);
      nl += pgemm(g2);
    }
    const float *fshift = fh_ada + (size_t)depth * 3 * C, *fscale = fshift + C;
    PK("layernorm", 0, 0, layernorm(fh_x, C, fh_h, C, B, C, nullptr, nullptr, 1e-6f, fscale, fshift, ada_n, st, s_h.out()));
    GemmArgs gf;
    gf.A = fh_h; gf.lda = C; gf.W = final_linear.w; gf.ldw = C; gf.M = B; gf.N = ldim; gf.K = C;
    gf.epi.bias = final_linear.b; gf.epi.alpha = inv;
    gf.epi.residual = fh_cur; gf.epi.ldr = ldim; gf.epi.C = fh_cur; gf.epi.ldc = ldim;
    gf.epi.Cs = s_cur.out();
    use_act(gf, s_h);
    nl += pgemm(gf);
  }
  CUDA_TRY(cudaGetLastError());
  return PTTS_OK;
}

// The device work of one AR frame for Bp rows (flow_lm.go:238-299): input
// gather / NaN->BOS / input_linear, 6 transformer layers, flow head, scatter of the
// new latent into the per-slot history.  Only fixed device pointers are used, so
// the sequence can be captured into a CUDA graph.
// Largest padded batch whose AR step is launched with programmatic dependent launch (PTTS_PDL_MAX_B to experiment).
static int pdl_max_b() {
  static const int v = getenv("PTTS_PDL_MAX_B") ? atoi(getenv("PTTS_PDL_MAX_B")) : 1 << 30;
  return v;
}

int Engine::step_body(int Bp, int max_keys, int lsd_steps, bool gather_frames, bool gather_noise,
                      bool timed, bool capturing) {
  // the ~79 short dependent kernels of a frame are chained with programmatic dependent
  // launch (launch.cuh); the edges survive graph capture
  static const bool no_pdl = getenv("PTTS_NO_PDL") != nullptr;
  struct PdlScope {
    PdlScope(bool on) { pdl_set(on); }
    ~PdlScope() { pdl_set(false); }
  } pdl_scope(!no_pdl && !prof_on && Bp <= pdl_max_b()); // measured: -3..9 % per frame up to B = 256, -5 % at 384..768, -1..3 % at 1024
  int64_t &nl = times.kernel_launches;
  if (Bp == 1 && fused_on && !prof_on) {
    // the batch-1 latency path: the whole frame is one persistent cooperative kernel
    float *tv = nullptr;
    int rc = get_tvec(lsd_steps, &tv);
    if (rc)
      return rc;
    const int n =
        fused_ring
            ? step_ring_launch(d_fused, ring_grid, ring_smem, ring_slots, d_row_slot, d_row_pos, d_off_a, d_off_b, d_off_c,
                               gather_frames ? 1 : 0, gather_noise ? 1 : 0, tv, lsd_steps, d_fused_bar, st)
            : step_fused_launch(d_fused, fused_grid, fused_smem, d_row_slot, d_row_pos, d_off_a, d_off_b, d_off_c,
                                gather_frames ? 1 : 0, gather_noise ? 1 : 0, tv, lsd_steps, d_fused_bar, fused_ll, st);
    if (n < 0) {
      cudaGetLastError();
      PTTS_FAIL(PTTS_ECUDA, "cuda: cooperative launch of the fused frame kernel failed");
    }
    nl += n;
    return PTTS_OK;
  }
  if (gather_frames)
    PK("gather_scatter", 0, 0, gather_rows(d_hist, d_off_a, fh_frames, Bp, ldim, st));
  if (gather_noise)
    PK("gather_scatter", 0, 0, gather_rows(d_noise, d_off_c, fh_noise, Bp, ldim, st));
  PK("elementwise", 0, 0, nan_to_bos(fh_frames, bos, fh_in, Bp, ldim, st, s_in.out()));
  {
    GemmArgs g;
    g.A = fh_in; g.lda = ldim; g.W = input_linear.w; g.ldw = ldim; g.M = Bp; g.N = D; g.K = ldim;
    g.epi.bias = input_linear.b; g.epi.C = x; g.epi.ldc = D;
    use_act(g, s_in);
    nl += pgemm(g);
  }
  int rc = flow_forward(Bp, max_keys, true);
  if (rc)
    return rc;
  // start of the flow head: a plain record when run eagerly, an external event-record node when
  // the body is being captured, so graph replays report flow_head_ms too
  if (timed)
    cudaEventRecord(ev[3], st);
  else if (capturing)
    cudaEventRecordWithFlags(ev[3], st, cudaEventRecordExternal);
  // one LSD step (the default): the whole head is one persistent cooperative kernel
  if (fh_on && lsd_steps == 1 && !prof_on && Bp >= tc_min_m)
    rc = flow_head_fused(Bp);
  else
    rc = flow_head(Bp, lsd_steps);
  if (rc)
    return rc;
  PK("gather_scatter", 0, 0, scatter_rows(fh_cur, d_hist, d_off_b, Bp, ldim, st));
  return PTTS_OK;
}

// One AR frame for B sequences.
//   frames_in_host: [B,ldim] or null (feed back hist[slot][frames]);
//   noise_host:     [B,ldim] or null (use d_noise[slot][frames]).
int Engine::step(const std::vector<Seq *> &seqs, const float *frames_in_host,
                 const float *noise_host, int lsd_steps, float *lat_out_host,
                 float *eos_out_host) {
  const int B = (int)seqs.size();
  if (B == 0)
    return PTTS_OK;
  if (B > max_batch)
    PTTS_FAIL(PTTS_EINVAL, "native-cuda: step batch %d exceeds max_batch %d", B, max_batch);
  if (lsd_steps <= 0)
    PTTS_FAIL(PTTS_EINVAL, "native: lsd decode steps must be >0");
  // rows are padded to a bucket so a handful of graphs covers every batch size;
  // pad rows have position -1 (skipped by RoPE/KV append/attention) and read /
  // write the dummy slot max_batch.
  int Bp = B <= 8 ? B : ((B + 15) / 16) * 16;
  if (Bp > max_batch)
    Bp = max_batch;
  int max_keys = 0;
  for (int i = 0; i < B; i++) {
    Seq *s = seqs[i];
    if (s->kv_len + 1 > s->kv_cap)
      PTTS_FAIL(PTTS_ENOMEM, "native-cuda: KV reservation exhausted (%d tokens)", s->kv_cap);
    if (s->frames >= max_frames)
      PTTS_FAIL(PTTS_ENOMEM, "native-cuda: frame budget exhausted (%d)", max_frames);
    h_row_slot[i] = s->slot;
    h_row_pos[i] = s->kv_len;
    h_off_a[i] = ((int64_t)s->slot * (max_frames + 1) + s->frames) * ldim;     // input frame
    h_off_b[i] = ((int64_t)s->slot * (max_frames + 1) + s->frames + 1) * ldim; // output frame
    h_off_c[i] = ((int64_t)s->slot * max_frames + s->frames) * ldim;           // noise row
    if (s->kv_len + 1 > max_keys)
      max_keys = s->kv_len + 1;
  }
  for (int i = B; i < Bp; i++) {
    h_row_slot[i] = 0;
    h_row_pos[i] = -1;
    h_off_a[i] = (int64_t)max_batch * (max_frames + 1) * ldim;
    h_off_b[i] = ((int64_t)max_batch * (max_frames + 1) + 1) * ldim;
    h_off_c[i] = (int64_t)max_batch * max_frames * ldim;
  }
  cudaEventRecord(ev[2], st);
  CUDA_TRY(cudaMemcpyAsync(d_row_slot, h_row_slot, (size_t)Bp * 4, cudaMemcpyHostToDevice, st));
  CUDA_TRY(cudaMemcpyAsync(d_row_pos, h_row_pos, (size_t)Bp * 4, cudaMemcpyHostToDevice, st));
  CUDA_TRY(cudaMemcpyAsync(d_off_b, h_off_b, (size_t)Bp * 8, cudaMemcpyHostToDevice, st));
  if (frames_in_host) {
    if (Bp > B)
      CUDA_TRY(cudaMemsetAsync(fh_frames, 0, (size_t)Bp * ldim * 4, st));
    CUDA_TRY(cudaMemcpyAsync(fh_frames, frames_in_host, (size_t)B * ldim * 4, cudaMemcpyHostToDevice, st));
  } else {
    CUDA_TRY(cudaMemcpyAsync(d_off_a, h_off_a, (size_t)Bp * 8, cudaMemcpyHostToDevice, st));
  }
  if (noise_host) {
    if (Bp > B)
      CUDA_TRY(cudaMemsetAsync(fh_noise, 0, (size_t)Bp * ldim * 4, st));
    CUDA_TRY(cudaMemcpyAsync(fh_noise, noise_host, (size_t)B * ldim * 4, cudaMemcpyHostToDevice, st));
  } else {
    CUDA_TRY(cudaMemcpyAsync(d_off_c, h_off_c, (size_t)Bp * 8, cudaMemcpyHostToDevice, st));
  }
  float *tv = nullptr;
  int rc = get_tvec(lsd_steps, &tv); // uploads on first use: must not happen inside a capture
  if (rc)
    return rc;
  const bool gf = frames_in_host == nullptr, gn = noise_host == nullptr;
  bool timed = true;
  if (graphs_on && !prof_on) {
    const uint64_t key = ((uint64_t)Bp << 32) | ((uint64_t)lsd_steps << 8) | (gf ? 2u : 0u) | (gn ? 1u : 0u);
    StepGraph &sg = step_graphs[key];
    sg.uses++;
    if (sg.exec) {
      CUDA_TRY(cudaGraphLaunch(sg.exec, st));
      times.kernel_launches += sg.launches;
    } else if (sg.uses >= 2) {
      const int64_t before = times.kernel_launches;
      cudaGraph_t graph = nullptr;
      CUDA_TRY(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
      // attention smem is sized for the per-sequence KV cap so one graph serves any length
      rc = step_body(Bp, cfg.max_seq_tokens, lsd_steps, gf, gn, false, true);
      cudaError_t ce = cudaStreamEndCapture(st, &graph);
      if (rc)
        return rc;
      if (ce != cudaSuccess)
        PTTS_FAIL(PTTS_ECUDA, "cuda: graph capture of the AR step failed: %s", cudaGetErrorString(ce));
      sg.launches = times.kernel_launches - before;
      ce = cudaGraphInstantiate(&sg.exec, graph, 0);
      cudaGraphDestroy(graph);
      if (ce != cudaSuccess) {
        sg.exec = nullptr;
        PTTS_FAIL(PTTS_ECUDA, "cuda: graph instantiate failed: %s", cudaGetErrorString(ce));
      }
      CUDA_TRY(cudaGraphLaunch(sg.exec, st));
    } else {
      rc = step_body(Bp, max_keys, lsd_steps, gf, gn, true);
      if (rc)
        return rc;
    }
  } else {
    rc = step_body(Bp, max_keys, lsd_steps, gf, gn, true);
    if (rc)
      return rc;
  }
  cudaEventRecord(ev[4], st);
  const bool used_fused = Bp == 1 && fused_on && !prof_on;
  if (used_fused) {
    timed = false; // no flow-head split inside the single kernel
    CUDA_TRY(cudaMemcpyAsync(h_fused_err, d_fused_err, sizeof(int), cudaMemcpyDeviceToHost, st));
  }
  const bool used_fh = fh_on && lsd_steps == 1 && !prof_on && Bp >= tc_min_m && !used_fused;
  if (used_fh)
    CUDA_TRY(cudaMemcpyAsync(h_fh_err, d_fh_err, sizeof(int), cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaMemcpyAsync(h_eos, fh_eos, (size_t)B * 4, cudaMemcpyDeviceToHost, st));
  if (lat_out_host)
    CUDA_TRY(cudaMemcpyAsync(h_lat, fh_cur, (size_t)B * ldim * 4, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  CUDA_TRY(cudaGetLastError());
  if (used_fused && fused_dbg && times.steps == 20) { // one frame's timeline of CTA 0 (tools)
    unsigned long long hd[512];
    cudaMemcpy(hd, fused_dbg, sizeof hd, cudaMemcpyDeviceToHost);
    fprintf(stderr, "fused frame timeline (ns after entry, CTA 0):");
    for (int i = 1; i < 512 && hd[i]; i++)
      fprintf(stderr, " %lld", (long long)(hd[i] - hd[0]));
    fprintf(stderr, "\n");
  }
  if (used_fused && *h_fused_err) {
    fused_on = false; // fall back to the multi-kernel path from here on
    cudaMemset(d_fused_err, 0, sizeof(int));
    PTTS_FAIL(PTTS_ECUDA, "native-cuda: fused frame kernel: grid barrier timed out");
  }
  if (used_fh && *h_fh_err) {
    fh_on = false; // the multi-kernel head from here on (graphs that embed the kernel are dropped)
    cudaMemset(d_fh_err, 0, sizeof(int));
    for (auto &kv : step_graphs)
      if (kv.second.exec)
        cudaGraphExecDestroy(kv.second.exec);
    step_graphs.clear();
    PTTS_FAIL(PTTS_ECUDA, "native-cuda: flow-head kernel: grid barrier timed out");
  }
  if (eos_out_host)
    memcpy(eos_out_host, h_eos, (size_t)B * 4);
  if (lat_out_host)
    memcpy(lat_out_host, h_lat, (size_t)B * ldim * 4);
  for (Seq *s : seqs) {
    s->kv_len++;
    s->frames++;
  }
  prof_collect();
  float ms = 0;
  cudaEventElapsedTime(&ms, ev[2], ev[4]);
  times.step_ms += ms;
  if (timed && cudaEventElapsedTime(&ms, ev[3], ev[4]) == cudaSuccess)
    times.flow_head_ms += ms;
  else
    cudaGetLastError();
  times.step_rows += B;
  times.steps += 1;
  debug_bufs["last_hidden"] = {fh_last, (size_t)B * D};
  debug_bufs["step_x"] = {x, (size_t)B * D};
  return PTTS_OK;
}

// ---------------------------------------------------------------------------
// Mimi decode of a ragged batch (model.go:141-167 + mimi.go:719-789).
// m_latin holds the gathered latent frames [total_frames, ldim].
// ---------------------------------------------------------------------------
int Engine::mimi_decode(std::vector<MimiSeg> &segs, int total_frames, bool d2h, PcmJob *job) {
  float *pcm = job ? m_pcm_ring[job->buf] : m_pcm;
  // Asynchronous groups (job) run on the Mimi stream beside the AR steps, everything else on
  // `st` after the Mimi stream has drained (the activation workspace is shared).  Profiled
  // passes keep one stream so the per-kernel event pairs do not measure contention.
  const bool side = job && overlap_on && !prof_on;
  if (!side)
    CUDA_TRY(cudaStreamSynchronize(st_mimi));
  struct LaneScope {
    Engine *e;
    LaneScope(Engine *e_, int l) : e(e_) { e->lane_ = l; }
    ~LaneScope() { e->lane_ = 0; }
  } lane_scope(this, side ? 1 : 0);
  cudaStream_t st = cur_st(); // shadows the member: every launch below goes to the lane's stream
  float *m_latin = job ? m_latin_ring[job->buf] : this->m_latin;
  int32_t *h_mrow_pos = job ? h_mrow_pos_ring[job->buf] : this->h_mrow_pos;
  int32_t *h_mrow_frame0 = job ? h_mrow_frame0_ring[job->buf] : this->h_mrow_frame0;
  uint8_t *h_mmask = job ? h_mmask_ring[job->buf] : this->h_mmask;
  if (job) {
    // the latent rows were gathered on the AR stream (decode_async)
    CUDA_TRY(cudaStreamWaitEvent(st, ev_lat_ready[job->buf], 0));
    CUDA_TRY(cudaMemcpyAsync(h_lat_ring[job->buf], m_latin, (size_t)total_frames * ldim * 4, cudaMemcpyDeviceToHost, st));
  }
  int64_t &nl = times.kernel_launches;
  const int P = mimi_pad;
  int Rb = 0;
  // a pass with stateful streaming segments carries absolute row positions (RoPE, attention window across
  // deliveries) and the per-row ring descriptors; the other segments of such a pass start at position 0
  bool stream = false;
  for (auto &sg : segs)
    stream = stream || sg.ring >= 0;
  stream = stream && job != nullptr;
  int32_t *h_apos = stream ? h_mrow_apos_ring[job->buf] : nullptr;
  int4 *h_seg = stream ? h_mrow_seg_ring[job->buf] : nullptr;
  for (auto &sg : segs) {
    sg.row0 = Rb;
    const int a0 = sg.ring >= 0 ? sg.abs_frame0 * up_stride : 0;
    for (int i = 0; i < P; i++) {
      h_mrow_pos[Rb + i] = -1;
      h_mrow_frame0[Rb + i] = sg.frame0;
      h_mmask[Rb + i] = 0;
      if (stream) {
        h_apos[Rb + i] = -1;
        h_seg[Rb + i] = make_int4(-1, 0, 0, 0);
      }
    }
    int n = sg.n_frames * up_stride;
    for (int i = 0; i < n; i++) {
      h_mrow_pos[Rb + P + i] = i + sg.pre * up_stride; // (a stream pass uses it for the upsampling taps only)
      h_mrow_frame0[Rb + P + i] = sg.frame0;
      h_mmask[Rb + P + i] = 1;
      if (stream) {
        h_apos[Rb + P + i] = a0 + i;
        h_seg[Rb + P + i] = make_int4(sg.ring, a0, Rb + P, 0);
      }
    }
    Rb += P + n;
  }
  if (Rb > mimi_rows_cap || total_frames > mimi_frames_cap)
    PTTS_FAIL(PTTS_ENOMEM, "native-cuda: mimi group (%d frames, %d rows) exceeds workspace", total_frames, Rb);
  cudaEventRecord(job ? ev_dec0[job->buf] : ev[5], st);
  CUDA_TRY(cudaMemcpyAsync(d_mrow_pos, h_mrow_pos, (size_t)Rb * 4, cudaMemcpyHostToDevice, st));
  CUDA_TRY(cudaMemcpyAsync(d_mrow_frame0, h_mrow_frame0, (size_t)Rb * 4, cudaMemcpyHostToDevice, st));
  CUDA_TRY(cudaMemcpyAsync(d_mmask, h_mmask, (size_t)Rb, cudaMemcpyHostToDevice, st));
  if (stream) {
    CUDA_TRY(cudaMemcpyAsync(d_mrow_apos, h_apos, (size_t)Rb * 4, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(d_mrow_seg, h_seg, (size_t)Rb * sizeof(int4), cudaMemcpyHostToDevice, st));
  }
  // ---- the kernels of the pass (captured once and replayed for small asynchronous passes) ----
  int rows_out = Rb;
  for (int s3 = 0; s3 < 3; s3++)
    rows_out *= up[s3].stride;
  const int spb = rows_out / Rb; // samples per base row
  auto run_kernels = [&]() -> int {
    // With the tensor-core path on, every GEMM / conv input is produced directly as split
    // bf16 planes (fetched by TMA) and its fp32 copy is not written.
    const bool sp_on = tc_on;
    auto f32 = [&](float *p) { return sp_on ? nullptr : p; };
    auto spo = [&](const ActSplit &a) { return sp_on ? a.out() : SplitOut(); };
    // latent -> mimi projection (folded denorm + 1x1 conv); m_latin is filled by the
    // caller (gather from the frame history, or host latents), split here
    if (sp_on && total_frames >= tc_min_m)
      PK("elementwise", 0, 0, copy_rows(m_latin, m_latin, total_frames, ldim, st, sm_latin.out()));
    {
      GemmArgs g;
      g.A = m_latin; g.lda = ldim; g.W = proj.w; g.ldw = ldim; g.M = total_frames; g.N = Md; g.K = ldim;
      g.epi.bias = proj.b; g.epi.C = m_lat; g.epi.ldc = Md;
      use_act(g, sm_latin);
      nl += pgemm(g);
    }
    PK("upsample_depthwise", 0, 0, upsample_depthwise(m_lat, up_w, m_x, Rb, Md, up_stride, up_k, d_mrow_pos, d_mrow_frame0, st));
    // decoder transformer (mimi.go:245-441)
    for (int l = 0; l < ML; l++) {
      DTLayer &ly = ml[l];
      PK("layernorm", 0, 0, layernorm(m_x, Md, f32(m_nrm), Md, Rb, Md, ly.n1.w, ly.n1.b, ly.n1.eps, nullptr, nullptr, 0, st, spo(sm_nrm)));
      GemmArgs g;
      g.A = f32(m_nrm); g.lda = Md; g.W = ly.in_proj.w; g.ldw = Md; g.M = Rb; g.N = 3 * Md; g.K = Md;
      g.epi.C = m_qkv; g.epi.ldc = 3 * Md;
      use_act(g, sm_nrm);
      nl += pgemm(g);
      PK("rope_inplace", 0, 0, rope_inplace(m_qkv, Rb, Md, MH, stream ? d_mrow_apos : d_mrow_pos, mrope_cos, mrope_sin, st));
      if (stream) {
        // stateful deliveries: this layer's K (rotated) and V rows join the requests' rings, the new rows attend
        // ring + buffer (the ring slots written here are never among the 249 older rows a row of the pass reads)
        const long loff = (long)l * ring_layer_stride;
        nl += kv_ring_append(m_qkv, Rb, Md, d_mrow_apos, d_mrow_seg, mimi_ring, ring_state_stride, loff, st);
        nl += attention_window_stream(m_qkv, 3 * Md, f32(m_att), Md, Rb, Md, MH, d_mrow_apos, d_mrow_seg, mimi_ring,
                                      ring_state_stride, loff, Mctx, st, spo(sm_att));
      } else if (prof_on) {
        // flops: every allowed (query, key) pair, 4 * head_dim per head.  bytes: what a flash-style
        // kernel must move — per block of 64 query rows the q rows, the UNIQUE K and V rows of the
        // block's windows (read once per block, not once per query) and the output rows.
        double keys = 0, rows_moved = 0;
        for (int r0 = 0; r0 < Rb; r0 += 64) {
          int lo = Rb, hi = -1;
          for (int r = r0; r < r0 + 64 && r < Rb; r++) {
            const int pos = h_mrow_pos[r];
            if (pos < 0)
              continue;
            const int n = pos + 1 < Mctx ? pos + 1 : Mctx;
            keys += n;
            if (r - n + 1 < lo)
              lo = r - n + 1;
            hi = r;
          }
          if (hi >= 0)
            rows_moved += 2.0 * (hi - lo + 1) + 2.0 * 64; // K + V window rows, q rows in, out rows
        }
        cudaEvent_t a, b;
        prof_begin(&a, &b);
        nl += attention_window(m_qkv, 3 * Md, f32(m_att), Md, Rb, Md, MH, d_mrow_pos, Mctx, st, spo(sm_att));
        prof_end(a, b, "attention_window", 4.0 * keys * Md, 4.0 * rows_moved * Md);
      } else
        nl += attention_window(m_qkv, 3 * Md, f32(m_att), Md, Rb, Md, MH, d_mrow_pos, Mctx, st, spo(sm_att));
      GemmArgs o;
      o.A = f32(m_att); o.lda = Md; o.W = ly.out_proj.w; o.ldw = Md; o.M = Rb; o.N = Md; o.K = Md;
      o.epi.colscale = ly.ls1; o.epi.residual = m_x; o.epi.ldr = Md; o.epi.C = m_x; o.epi.ldc = Md;
      use_act(o, sm_att);
      nl += pgemm(o);
      PK("layernorm", 0, 0, layernorm(m_x, Md, f32(m_nrm), Md, Rb, Md, ly.n2.w, ly.n2.b, ly.n2.eps, nullptr, nullptr, 0, st, spo(sm_nrm)));
      GemmArgs f1;
      f1.A = f32(m_nrm); f1.lda = Md; f1.W = ly.l1.w; f1.ldw = Md; f1.M = Rb; f1.N = MF; f1.K = Md;
      f1.epi.act = ACT_GELU; f1.epi.C = f32(m_ff); f1.epi.ldc = MF; f1.epi.Cs = spo(sm_ff);
      use_act(f1, sm_nrm);
      nl += pgemm(f1);
      GemmArgs f2;
      f2.A = f32(m_ff); f2.lda = MF; f2.W = ly.l2.w; f2.ldw = MF; f2.M = Rb; f2.N = Md; f2.K = MF;
      f2.epi.colscale = ly.ls2; f2.epi.residual = m_x; f2.epi.ldr = Md; f2.epi.C = m_x; f2.epi.ldc = Md;
      if (l == ML - 1) { // pad rows must be zero again: they are the conv left context
        f2.epi.rowmask = d_mmask;
        f2.epi.mask_div = 1;
        f2.epi.Cs = spo(sm_x); // the SEANet input
      }
      use_act(f2, sm_ff);
      nl += pgemm(f2);
    }
    debug_bufs["mimi_transformer"] = {m_x, (size_t)Rb * Md};
    // SEANet (mimi.go:745-787): convs as GEMMs over overlapping channels-last rows
    auto conv_gemm = [&](const DConvG &c, const float *in, const ActSplit &ins, int rows, int div, float *raw,
                         float *elu, const ActSplit *elus, const float *residual) {
      GemmArgs g;
      g.A = sp_on ? nullptr : in - (size_t)(c.taps - 1) * c.cin; g.lda = c.cin; g.W = c.w; g.ldw = c.K;
      g.M = rows; g.N = c.N; g.K = c.K;
      g.epi.bias = c.b; g.epi.C = raw; g.epi.ldc = c.N;
      g.epi.C2 = elu; g.epi.ldc2 = c.N;
      if (elus && sp_on) {
        // a stride-s transposed conv writes s output rows per input row: the split planes of
        // its output have c.N / s channels per row, the same memory seen as [rows, c.N]
        g.epi.C2s = elus->out();
        g.epi.C2s.ld = c.N;
      }
      g.epi.residual = residual; g.epi.ldr = c.N;
      g.epi.rowmask = d_mmask; g.epi.mask_div = div;
      if (sp_on) {
        if (c.taps > 1) {
          g.amap = c.amap;
          g.a_lo_row = (int)ins.plane_rows;
          g.a_cin = c.cin;
        } else
          use_act(g, ins);
      }
      nl += pgemm(g);
    };
    conv_gemm(init_conv, m_x, sm_x, Rb, 1, nullptr, f32(m_y0e), &sm_y0e, nullptr);
    const float *cur = m_y0e;
    const ActSplit *curs = &sm_y0e;
    bool cur_alt = false; // the previous stage's fused block wrote sm_yf instead of sm_ye
    int rows = Rb, div = 1;
    for (int s = 0; s < 3; s++) {
      if (cur_alt) {
        DConvG u = up[s];
        u.amap = up_alt_amap[s];
        conv_gemm(u, cur, *curs, rows, div, m_y[s], f32(m_ye[s]), &sm_ye[s], nullptr);
      } else
        conv_gemm(up[s], cur, *curs, rows, div, m_y[s], f32(m_ye[s]), &sm_ye[s], nullptr); // rows*stride outputs
      cur_alt = false;
      rows *= up[s].stride;
      div *= up[s].stride;
      // elu(x + conv2(elu(conv1(elu(x))))); the last stage feeds the fp32 final conv
      const bool last = s == 2;
      bool fused = false;
      static const bool no_rb = getenv("PTTS_NO_FUSED_RESBLOCK") != nullptr;
      if (sp_on && !no_rb && res1[s].amap && res2[s].taps == 1 && (last || (sm_yf[s].p && up_alt_amap[s + 1]))) {
        auto w1 = tc_map.find(res1[s].w), w2 = tc_map.find(res2[s].w);
        if (w1 != tc_map.end() && w2 != tc_map.end()) {
          cudaEvent_t pa, pb;
          if (prof_on)
            prof_begin(&pa, &pb);
          const int n = resblock_fused(*w1->second, *w2->second, res1[s].amap, (int)sm_ye[s].plane_rows, rows, res1[s].cin,
                                       res1[s].cout, res1[s].taps, res1[s].b, res2[s].b, m_y[s], d_mmask, div,
                                       last ? m_ye[s] : nullptr, last ? SplitOut() : sm_yf[s].out(), st);
          if (prof_on) {
            const double c = res1[s].cin, h = res1[s].cout;
            prof_end(pa, pb, n ? "resblock_fused" : "none", 2.0 * rows * (3 * c * h + h * c), 4.0 * rows * 3 * c,
                     s == 2 ? "shape/resblock3" : "shape/resblock2");
          }
          nl += n;
          fused = n > 0;
        }
      }
      const ActSplit *next_src = &sm_ye[s];
      if (fused && !last) {
        next_src = &sm_yf[s];
        cur_alt = true;
      }
      if (!fused) {
        conv_gemm(res1[s], m_ye[s], sm_ye[s], rows, div, nullptr, f32(m_he[s]), &sm_he[s], nullptr);
        conv_gemm(res2[s], m_he[s], sm_he[s], rows, div, nullptr, (last || !sp_on) ? m_ye[s] : nullptr,
                  last ? nullptr : &sm_ye[s], m_y[s]);
      }
      cur = m_ye[s];
      curs = next_src;
    }
    PK("final_conv", 0, 0, final_conv(cur, rows, final_cin, final_k, final_w, final_b, d_mmask, div, pcm, st));
    return PTTS_OK;
  };
  StepGraph *mg = nullptr;
  if (side && mimi_graphs_on && graphs_on && total_frames <= 16 && job) {
    const uint64_t key = ((uint64_t)Rb << 40) | ((uint64_t)total_frames << 24) | ((uint64_t)segs.size() << 8) |
                         ((uint64_t)job->buf << 1) | (stream ? 1u : 0u);
    auto it = mimi_graphs.find(key);
    if (it != mimi_graphs.end())
      mg = &it->second;
    else if (mimi_graphs.size() < 256) // shapes beyond that run kernel by kernel
      mg = &mimi_graphs[key];
    if (mg)
      mg->uses++;
  }
  if (mg && mg->exec) {
    CUDA_TRY(cudaGraphLaunch(mg->exec, st));
    nl += mg->launches;
  } else if (mg && mg->uses >= 2) {
    const int64_t before = nl;
    cudaGraph_t graph = nullptr;
    CUDA_TRY(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
    const int rc = run_kernels();
    cudaError_t ce = cudaStreamEndCapture(st, &graph);
    if (rc)
      return rc;
    if (ce != cudaSuccess)
      PTTS_FAIL(PTTS_ECUDA, "cuda: graph capture of a Mimi pass failed: %s", cudaGetErrorString(ce));
    mg->launches = nl - before;
    ce = cudaGraphInstantiate(&mg->exec, graph, 0);
    cudaGraphDestroy(graph);
    if (ce != cudaSuccess) {
      mg->exec = nullptr;
      PTTS_FAIL(PTTS_ECUDA, "cuda: graph instantiate of a Mimi pass failed: %s", cudaGetErrorString(ce));
    }
    CUDA_TRY(cudaGraphLaunch(mg->exec, st));
  } else {
    const int rc = run_kernels();
    if (rc)
      return rc;
  }
  const int rows = rows_out;
  debug_bufs["mimi_pcm"] = {pcm, (size_t)rows};
  times.mimi_frames += total_frames;
  if (job) {
    // asynchronous return: the copy stream takes over once the PCM is final; the
    // finisher thread waits on ev_copied, fills the callers' buffers and completes them
    const int k = job->buf;
    if (job->copy_s16)
      nl += pcm_to_s16(pcm, m_s16_ring[k], (size_t)rows, st);
    cudaEventRecord(ev_pcm_ready[k], st);
    CUDA_TRY(cudaStreamWaitEvent(st_copy, ev_pcm_ready[k], 0));
    cudaEventRecord(ev_copy0[k], st_copy);
    if (job->copy)
      CUDA_TRY(cudaMemcpyAsync(h_pcm_ring[k], pcm, (size_t)rows * sizeof(float), cudaMemcpyDeviceToHost, st_copy));
    if (job->copy_s16)
      CUDA_TRY(cudaMemcpyAsync(h_s16_ring[k], m_s16_ring[k], (size_t)rows * sizeof(int16_t), cudaMemcpyDeviceToHost,
                               st_copy));
    cudaEventRecord(ev_copied[k], st_copy);
    for (size_t i = 0; i < segs.size(); i++)
      job->segs[i].pcm_off = (size_t)(segs[i].row0 + P) * spb;
    CUDA_TRY(cudaGetLastError());
    if (prof_on) {
      CUDA_TRY(cudaStreamSynchronize(st));
      prof_collect();
    }
    return PTTS_OK;
  }
  cudaEventRecord(ev[6], st);
  if (d2h) {
    for (auto &sg : segs)
      if (sg.host_out)
        CUDA_TRY(cudaMemcpyAsync(sg.host_out, pcm + (size_t)(sg.row0 + P) * spb,
                                 (size_t)sg.n_frames * up_stride * spb * 4, cudaMemcpyDeviceToHost, st));
  }
  cudaEventRecord(ev[7], st);
  CUDA_TRY(cudaStreamSynchronize(st));
  CUDA_TRY(cudaGetLastError());
  prof_collect();
  float ms = 0;
  cudaEventElapsedTime(&ms, ev[5], ev[6]);
  times.mimi_ms += ms;
  cudaEventElapsedTime(&ms, ev[6], ev[7]);
  times.d2h_ms += ms;
  return PTTS_OK;
}

// Scheduler path: gather + Mimi-decode the finished requests in workspace-sized groups
// and hand each group to the finisher thread (PCM D2H and request completion overlap the
// following AR steps).  The sequences can be released as soon as this returns: every read
// of their frame history is already ordered on the stream.
int Engine::decode_async(const std::vector<DecodeItem> &items) {
  ensure_finisher();
  size_t i = 0;
  const int max_segs = (mimi_rows_cap - up_stride * mimi_frames_cap) / mimi_pad;
  while (i < items.size()) {
    PcmJob job;
    std::vector<MimiSeg> segs;
    int tf = 0;
    const int k = pcm_next;
    {
      std::unique_lock<std::mutex> lk(fin_mu);
      fin_cv.wait(lk, [&] { return !buf_busy[k]; });
      buf_busy[k] = true;
    }
    pcm_next = (pcm_next + 1) % NPCM;
    job.buf = k;
    while (i < items.size() && (int)segs.size() < max_segs) {
      const DecodeItem &it = items[i];
      const std::shared_ptr<Request> &r = it.r;
      const int n = it.count;
      if (n <= 0 || n > mimi_frames_cap) { // cannot happen (finalize raises the workspace to max_frames): fail it alone
        char msg[128];
        snprintf(msg, sizeof msg, "native-cuda: %d frames exceed mimi_group_frames %d", n, mimi_frames_cap);
        complete_request(r, PTTS_ENOMEM, msg);
        i++;
        continue;
      }
      if (tf + n + it.pre > mimi_frames_cap)
        break;
      MimiSeg sg;
      sg.n_frames = n;
      sg.frame0 = tf;
      sg.row0 = 0;
      sg.host_out = nullptr;
      sg.ring = it.ring;
      sg.abs_frame0 = it.first;
      sg.pre = it.pre;
      for (int f = 0; f < it.pre; f++, tf++)
        h_moff_ring[k][tf] = ((int64_t)r->seq->slot * (max_frames + 1) + it.first - it.pre + f + 1) * ldim;
      for (int f = 0; f < n; f++)
        h_moff_ring[k][tf + f] = ((int64_t)r->seq->slot * (max_frames + 1) + it.first + f + 1) * ldim;
      PcmJob::Seg js;
      js.r = r;
      js.lat_off = (size_t)tf * ldim;
      js.n_frames = n;
      js.skip = it.skip;
      js.out_frame0 = it.first + it.skip;
      js.final = it.final;
      js.stream_seq = r->stream_frames > 0 ? ++r->stream_seq : 0;
      job.segs.push_back(js);
      if (!(r->cfg.flags & PTTS_REQ_NO_PCM_D2H) && r->pcm_out) {
        if (r->cfg.flags & PTTS_REQ_PCM16)
          job.copy_s16 = true;
        else
          job.copy = true;
      }
      tf += n;
      segs.push_back(sg);
      i++;
    }
    if (segs.empty()) {
      std::unique_lock<std::mutex> lk(fin_mu);
      buf_busy[k] = false;
      continue;
    }
    int rc = PTTS_OK;
    // The latent rows leave the per-slot history on the AR stream (ordered after the step that
    // wrote them and before any later sequence reuses the slot); the decode itself follows on
    // the Mimi stream.  Host staging is per ring slot, so nothing here waits for the GPU.
    if (cudaMemcpyAsync(d_moff, h_moff_ring[k], (size_t)tf * 8, cudaMemcpyHostToDevice, st) != cudaSuccess)
      rc = PTTS_ECUDA;
    if (rc == PTTS_OK) {
      times.kernel_launches += gather_rows(d_hist, d_moff, m_latin_ring[k], tf, ldim, st);
      cudaEventRecord(ev_lat_ready[k], st);
      rc = mimi_decode(segs, tf, false, &job);
    }
    if (rc) {
      job.rc = rc;
      job.msg = last_error();
      // make the events valid for the finisher even though nothing ran
      cudaEventRecord(ev_lat_ready[k], st);
      cudaEventRecord(ev_dec0[k], st);
      cudaEventRecord(ev_pcm_ready[k], st);
      cudaStreamWaitEvent(st_copy, ev_pcm_ready[k], 0);
      cudaEventRecord(ev_copy0[k], st_copy);
      cudaEventRecord(ev_copied[k], st_copy);
    }
    {
      std::unique_lock<std::mutex> lk(fin_mu);
      fin_q.push_back(std::move(job));
    }
    fin_cv.notify_all();
  }
  return PTTS_OK;
}

// Decode frame spans of live sequences, grouped to fit the workspace.
int Engine::decode_seqs(const std::vector<Seq *> &seqs, const std::vector<int> &first,
                        const std::vector<int> &count, const std::vector<float *> &outs,
                        bool d2h) {
  size_t i = 0;
  const int max_segs = (mimi_rows_cap - up_stride * mimi_frames_cap) / mimi_pad;
  while (i < seqs.size()) {
    std::vector<MimiSeg> segs;
    int tf = 0;
    while (i < seqs.size() && (int)segs.size() < max_segs) {
      int n = count[i];
      if (n <= 0 || first[i] < 0 || first[i] + n > seqs[i]->frames)
        PTTS_FAIL(PTTS_EINVAL, "native-cuda: decode span [%d,%d) outside generated frames %d",
                  first[i], first[i] + n, seqs[i]->frames);
      if (n > mimi_frames_cap)
        PTTS_FAIL(PTTS_ENOMEM, "native-cuda: %d frames exceed mimi_group_frames %d", n, mimi_frames_cap);
      if (tf + n > mimi_frames_cap)
        break;
      MimiSeg sg;
      sg.n_frames = n;
      sg.frame0 = tf;
      sg.row0 = 0;
      sg.host_out = outs[i];
      for (int f = 0; f < n; f++)
        h_moff[tf + f] = ((int64_t)seqs[i]->slot * (max_frames + 1) + first[i] + f + 1) * ldim;
      tf += n;
      segs.push_back(sg);
      i++;
    }
    CUDA_TRY(cudaMemcpyAsync(d_moff, h_moff, (size_t)tf * 8, cudaMemcpyHostToDevice, st));
    times.kernel_launches += gather_rows(d_hist, d_moff, m_latin, tf, ldim, st);
    int rc = mimi_decode(segs, tf, d2h);
    if (rc)
      return rc;
  }
  return PTTS_OK;
}

} // namespace ptts
