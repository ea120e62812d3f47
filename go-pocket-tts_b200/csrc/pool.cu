// pool.cu — one process, several GPUs: a pool of engines (one per device, each with a full
// replica of the weights, its own paged KV pool and continuous-batching scheduler thread) behind
// the same submit / wait calls.
//
// Replaces the reference server's worker semaphore (internal/server/server.go:114,402-421) for a
// multi-GPU box: chunks are independent (internal/tts/service.go:138-153 generates every <=50
// token chunk with fresh state), so a submitted chunk goes to the least-loaded device and nothing
// crosses GPUs — no collective, no peer traffic.  Load = frames still to generate
// + prompt tokens still to prefill over the device's unfinished requests; ties rotate.
//
// A pool ticket is (engine index << 48) | engine ticket.
#include "engine.cuh"
#include <cstring>

namespace ptts {
const char *last_error();
}
using namespace ptts;

struct ptts_pool {
  std::vector<ptts_engine *> engines;
  std::atomic<unsigned> rr{0};
  std::vector<std::atomic<long>> dispatched; // requests sent to each engine (diagnostics)
  explicit ptts_pool(size_t n) : dispatched(n) {
    for (auto &d : dispatched)
      d.store(0);
  }
};

static const int64_t kTicketMask = (int64_t(1) << 48) - 1;

static int pool_split(ptts_pool *p, int64_t ticket, ptts_engine **e, int64_t *t) {
  if (!p)
    PTTS_FAIL(PTTS_EINVAL, "native-cuda runtime unavailable");
  const int64_t idx = ticket >> 48;
  if (ticket <= 0 || idx < 0 || idx >= (int64_t)p->engines.size())
    PTTS_FAIL(PTTS_EINVAL, "native-cuda: unknown pool ticket %lld", (long long)ticket);
  *e = p->engines[(size_t)idx];
  *t = ticket & kTicketMask;
  return PTTS_OK;
}

extern "C" {

int ptts_pool_create(const ptts_config *cfg, const int32_t *devices, int32_t n_devices, ptts_pool **out) {
  if (!out)
    PTTS_FAIL(PTTS_EINVAL, "native-cuda: ptts_pool_create: out is NULL");
  *out = nullptr;
  std::vector<int32_t> devs;
  if (devices && n_devices > 0)
    devs.assign(devices, devices + n_devices);
  else {
    const int n = ptts_device_count(); // default: every visible GPU
    for (int i = 0; i < n; i++)
      devs.push_back(i);
  }
  if (devs.empty())
    PTTS_FAIL(PTTS_ECUDA, "native-cuda runtime unavailable: no CUDA device");
  if (devs.size() > 64)
    PTTS_FAIL(PTTS_EINVAL, "native-cuda: pool of %zu devices (max 64)", devs.size());
  ptts_pool *p = new ptts_pool(devs.size());
  for (int32_t d : devs) {
    ptts_config c{};
    if (cfg)
      c = *cfg;
    c.device = d;
    ptts_engine *e = nullptr;
    int rc = ptts_create(&c, &e);
    if (rc) {
      for (ptts_engine *x : p->engines)
        ptts_destroy(x);
      delete p;
      return rc;
    }
    p->engines.push_back(e);
  }
  *out = p;
  return PTTS_OK;
}

void ptts_pool_destroy(ptts_pool *p) {
  if (!p)
    return;
  for (ptts_engine *e : p->engines)
    ptts_destroy(e);
  delete p;
}

int ptts_pool_size(ptts_pool *p) { return p ? (int)p->engines.size() : 0; }

int ptts_pool_engine(ptts_pool *p, int32_t index, ptts_engine **out) {
  if (!p || !out || index < 0 || index >= (int32_t)p->engines.size())
    PTTS_FAIL(PTTS_EINVAL, "native-cuda: pool engine index out of range");
  *out = p->engines[(size_t)index];
  return PTTS_OK;
}

// Weights: every engine gets its own replica (the host tensor is read once per device).
int ptts_pool_load_tensor(ptts_pool *p, const char *key, const int64_t *shape, int32_t rank, int32_t dtype,
                          const void *host_ptr) {
  if (!p)
    PTTS_FAIL(PTTS_EINVAL, "native-cuda runtime unavailable");
  for (ptts_engine *e : p->engines) {
    int rc = ptts_load_tensor(e, key, shape, rank, dtype, host_ptr);
    if (rc)
      return rc;
  }
  return PTTS_OK;
}

int ptts_pool_load_safetensors(ptts_pool *p, const char *path) {
  if (!p)
    PTTS_FAIL(PTTS_EINVAL, "native-cuda runtime unavailable");
  for (ptts_engine *e : p->engines) {
    int rc = ptts_load_safetensors(e, path);
    if (rc)
      return rc;
  }
  return PTTS_OK;
}

int ptts_pool_finalize_weights(ptts_pool *p) {
  if (!p)
    PTTS_FAIL(PTTS_EINVAL, "native-cuda runtime unavailable");
  // the uploads of different devices are independent: one thread per engine
  std::vector<int> rcs(p->engines.size(), 0);
  std::vector<std::string> msgs(p->engines.size());
  std::vector<std::thread> ths;
  for (size_t i = 0; i < p->engines.size(); i++)
    ths.emplace_back([&, i] {
      rcs[i] = ptts_finalize_weights(p->engines[i]);
      if (rcs[i])
        msgs[i] = ptts::last_error();
    });
  for (auto &t : ths)
    t.join();
  for (size_t i = 0; i < rcs.size(); i++)
    if (rcs[i])
      PTTS_FAIL(rcs[i], "device %d: %s", p->engines[i]->impl.device, msgs[i].c_str());
  return PTTS_OK;
}

// Least-loaded dispatch.  Returns the engine index through *engine_index (may be NULL).
int ptts_pool_submit(ptts_pool *p, const int64_t *tokens, int32_t n_tokens, const float *voice_emb,
                     int32_t voice_frames, const ptts_voice_kv *voice_kv, const ptts_gen_cfg *cfg, const float *noise,
                     int32_t noise_rows, float *pcm_out, int64_t pcm_cap, int64_t *ticket, int32_t *engine_index) {
  if (!p || !ticket)
    PTTS_FAIL(PTTS_EINVAL, "native-cuda runtime unavailable");
  const size_t n = p->engines.size();
  const unsigned start = p->rr.fetch_add(1);
  size_t best = start % n;
  long best_load = p->engines[best]->impl.inflight_cost.load();
  for (size_t k = 1; k < n; k++) {
    const size_t i = (start + k) % n;
    const long l = p->engines[i]->impl.inflight_cost.load();
    if (l < best_load) {
      best = i;
      best_load = l;
    }
  }
  int64_t t = 0;
  int rc = ptts_submit(p->engines[best], tokens, n_tokens, voice_emb, voice_frames, voice_kv, cfg, noise, noise_rows,
                       pcm_out, pcm_cap, &t);
  if (rc)
    return rc;
  p->dispatched[best].fetch_add(1);
  *ticket = ((int64_t)best << 48) | t;
  if (engine_index)
    *engine_index = (int32_t)best;
  return PTTS_OK;
}

int ptts_pool_wait(ptts_pool *p, int64_t ticket, ptts_result *result) {
  ptts_engine *e;
  int64_t t;
  int rc = pool_split(p, ticket, &e, &t);
  return rc ? rc : ptts_wait(e, t, result);
}

int ptts_pool_cancel(ptts_pool *p, int64_t ticket) {
  ptts_engine *e;
  int64_t t;
  int rc = pool_split(p, ticket, &e, &t);
  return rc ? rc : ptts_cancel(e, t);
}

int ptts_pool_ticket_error(ptts_pool *p, int64_t ticket, char *buf, int32_t cap) {
  ptts_engine *e;
  int64_t t;
  int rc = pool_split(p, ticket, &e, &t);
  return rc ? rc : ptts_ticket_error(e, t, buf, cap);
}

int ptts_pool_ticket_latents(ptts_pool *p, int64_t ticket, float *latents, float *eos_logits, int32_t cap_frames) {
  ptts_engine *e;
  int64_t t;
  int rc = pool_split(p, ticket, &e, &t);
  return rc ? rc : ptts_ticket_latents(e, t, latents, eos_logits, cap_frames);
}

int ptts_pool_ticket_read(ptts_pool *p, int64_t ticket, int32_t have_frames, int32_t timeout_ms, int32_t *frames_ready,
                          int32_t *done) {
  ptts_engine *e;
  int64_t t;
  int rc = pool_split(p, ticket, &e, &t);
  return rc ? rc : ptts_ticket_read(e, t, have_frames, timeout_ms, frames_ready, done);
}

int ptts_pool_ticket_release(ptts_pool *p, int64_t ticket) {
  ptts_engine *e;
  int64_t t;
  int rc = pool_split(p, ticket, &e, &t);
  return rc ? rc : ptts_ticket_release(e, t);
}

// Blocking form: tts.Runtime.GenerateAudio on a multi-GPU box.
int ptts_pool_generate_audio(ptts_pool *p, const int64_t *tokens, int32_t n_tokens, const float *voice_emb,
                             int32_t voice_frames, const ptts_voice_kv *voice_kv, const ptts_gen_cfg *cfg,
                             const float *noise, int32_t noise_rows, float *pcm_out, int64_t pcm_cap,
                             ptts_result *result) {
  int64_t t = 0;
  int rc = ptts_pool_submit(p, tokens, n_tokens, voice_emb, voice_frames, voice_kv, cfg, noise, noise_rows, pcm_out,
                            pcm_cap, &t, nullptr);
  if (rc)
    return rc;
  rc = ptts_pool_wait(p, t, result);
  ptts_pool_ticket_release(p, t);
  return rc;
}

// Per-engine load as the dispatcher sees it + how many requests each engine has been given.
int ptts_pool_load(ptts_pool *p, int64_t *inflight_cost, int32_t *inflight_requests, int64_t *dispatched, int32_t cap) {
  if (!p)
    PTTS_FAIL(PTTS_EINVAL, "native-cuda runtime unavailable");
  const int n = (int)p->engines.size() < cap ? (int)p->engines.size() : cap;
  for (int i = 0; i < n; i++) {
    if (inflight_cost)
      inflight_cost[i] = p->engines[(size_t)i]->impl.inflight_cost.load();
    if (inflight_requests)
      inflight_requests[i] = p->engines[(size_t)i]->impl.inflight_requests.load();
    if (dispatched)
      dispatched[i] = p->dispatched[(size_t)i].load();
  }
  return n;
}

} // extern "C"
