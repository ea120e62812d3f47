// scheduler.cu — continuous-batching scheduler + the extern "C" boundary.
//
// Replaces, behind the same blocking GenerateAudio contract, the reference's
// per-request B=1 loop whose AR step is serialised by rngMu
// (internal/tts/runtime_native_safetensors.go:155-198) and the server's
// semaphore worker pool (internal/server/server.go:114,402-421): every
// in-flight request owns a sequence slot; each scheduler iteration admits and
// prefills new requests, runs ONE batched AR frame for all live sequences,
// applies the reference's EOS-countdown rule per request on the host, and
// batch-decodes finished chunks through Mimi.
#include "engine.cuh"
#include <cmath>
#include <cstring>

namespace ptts {

const char *last_error();

static int resolve_max_steps(const ptts_gen_cfg &c, int ntok) {
  // runtime_native_safetensors.go:61-67 + internal/text/prepare.go:38-48
  int ms = c.max_steps;
  if (ms <= 0)
    ms = c.estimated_max_steps;
  if (ms <= 0) {
    int n = ntok < 0 ? 0 : ntok;
    ms = (int)std::ceil(((double)n / 3.0 + 2.0) * 12.5);
  }
  return ms;
}

void complete_request(const std::shared_ptr<Request> &r, int status, const char *msg) {
  std::unique_lock<std::mutex> lk(r->mu);
  if (r->done)
    return;
  if (r->eng) {
    r->eng->inflight_cost.fetch_sub(r->cost);
    r->eng->inflight_requests.fetch_sub(1);
  }
  r->res.status = status;
  if (msg)
    r->err = msg;
  auto now = std::chrono::steady_clock::now();
  r->res.total_ms = std::chrono::duration<double, std::milli>(now - r->t_submit).count();
  r->done = true;
  r->cv.notify_all();
}

void Engine::ensure_sched() {
  std::unique_lock<std::mutex> lk(q_mu);
  if (!sched_running) {
    sched_running = true;
    sched = std::thread([this] { sched_loop(); });
  }
}

void Engine::ensure_finisher() {
  std::unique_lock<std::mutex> lk(fin_mu);
  if (!fin_running) {
    fin_running = true;
    for (int i = 0; i < NFIN; i++)
      finishers.emplace_back([this] { finisher_loop(); });
  }
}

// Completes decoded requests: waits for the group's PCM to land in the pinned ring,
// copies each request's samples / latents into its own buffers, signals the waiter.
void Engine::finisher_loop() {
  cudaSetDevice(device);
  for (;;) {
    PcmJob job;
    {
      std::unique_lock<std::mutex> lk(fin_mu);
      fin_cv.wait(lk, [&] { return fin_stop || !fin_q.empty(); });
      if (fin_q.empty())
        break; // fin_stop and drained
      job = std::move(fin_q.front());
      fin_q.pop_front();
    }
    const int k = job.buf;
    cudaError_t ce = cudaEventSynchronize(ev_copied[k]);
    if (ce != cudaSuccess && job.rc == PTTS_OK) {
      job.rc = PTTS_ECUDA;
      job.msg = std::string("cuda: mimi decode failed: ") + cudaGetErrorString(ce);
    }
    float ms = 0;
    if (job.rc == PTTS_OK && cudaEventElapsedTime(&ms, ev_dec0[k], ev_pcm_ready[k]) == cudaSuccess) {
      float ms2 = 0; // PCM device -> pinned host on the copy stream
      if (cudaEventElapsedTime(&ms2, ev_copy0[k], ev_copied[k]) != cudaSuccess)
        ms2 = 0;
      std::unique_lock<std::mutex> lk(fin_mu);
      fin_mimi_ms += ms;
      fin_d2h_ms += ms2;
    }
    const int spf = info.samples_per_frame;
    for (auto &sg : job.segs) {
      Request &r = *sg.r;
      const int n_out = sg.n_frames - sg.skip; // the leading `skip` frames were decoded for context only
      // Deliveries of a streaming request are applied in issue order, and under the request's lock:
      // a cancelled / failed request is complete, its caller may already have released pcm_out.
      std::unique_lock<std::mutex> slk(r.mu, std::defer_lock);
      if (sg.stream_seq > 0) {
        slk.lock();
        r.cv.wait(slk, [&] { return r.stream_applied == sg.stream_seq - 1 || r.done; });
      }
      if (job.rc == PTTS_OK && n_out > 0 && !(sg.stream_seq > 0 && r.done)) {
        const float *lat = h_lat_ring[k] + sg.lat_off + (size_t)sg.skip * ldim;
        if (sg.stream_seq > 0)
          r.latents.insert(r.latents.end(), lat, lat + (size_t)n_out * ldim);
        else
          r.latents.assign(lat, lat + (size_t)n_out * ldim);
        if (!(r.cfg.flags & PTTS_REQ_NO_PCM_D2H) && r.pcm_out) {
          const size_t src = sg.pcm_off + (size_t)sg.skip * spf, dst = (size_t)sg.out_frame0 * spf;
          if (r.cfg.flags & PTTS_REQ_PCM16)
            memcpy(reinterpret_cast<int16_t *>(r.pcm_out) + dst, h_s16_ring[k] + src, (size_t)n_out * spf * sizeof(int16_t));
          else
            memcpy(r.pcm_out + dst, h_pcm_ring[k] + src, (size_t)n_out * spf * sizeof(float));
        }
      }
      if (sg.stream_seq > 0) {
        r.stream_applied = sg.stream_seq;
        if (job.rc == PTTS_OK && !r.done)
          r.frames_ready = sg.out_frame0 + n_out;
        r.cv.notify_all();
        slk.unlock();
      }
      if (sg.final || job.rc)
        complete_request(sg.r, job.rc, job.rc ? job.msg.c_str() : nullptr);
    }
    {
      std::unique_lock<std::mutex> lk(fin_mu);
      buf_busy[k] = false;
    }
    fin_cv.notify_all();
  }
}

void Engine::release_stream_state(Request &r) {
  if (r.stream_state >= 0) {
    free_stream_states.push_back(r.stream_state);
    r.stream_state = -1;
  }
}

void Engine::sched_loop() {
  cudaSetDevice(device);
  std::vector<std::shared_ptr<Request>> active, pending;
  int pending_frames = 0;
  long iter = 0;
  for (;;) {
    std::vector<std::shared_ptr<Request>> candidates;
    {
      std::unique_lock<std::mutex> lk(q_mu);
      if (active.empty() && pending.empty())
        q_cv.wait(lk, [&] { return sched_stop || !queue.empty(); });
      if (sched_stop) {
        // Close() with requests in flight: every waiter is released (queued, live and
        // finished-but-undecoded alike); their sequences go back to the pool.
        for (auto &r : queue)
          complete_request(r, PTTS_ECANCELED, "native-cuda: engine destroyed");
        queue.clear();
        lk.unlock();
        std::unique_lock<std::mutex> gl(gpu_mu);
        for (auto *v : {&active, &pending})
          for (auto &r : *v) {
            if (r->seq) {
              seq_free(r->seq);
              delete reinterpret_cast<ptts_seq *>(r->seq);
              r->seq = nullptr;
            }
            complete_request(r, PTTS_ECANCELED, "native-cuda: engine destroyed");
          }
        active.clear();
        pending.clear();
        break;
      }
      // take as many queued requests as there could be room for; resources are
      // claimed below under gpu_mu, leftovers are pushed back in order.
      while (!queue.empty() && (int)(active.size() + pending.size() + candidates.size()) < max_batch) {
        candidates.push_back(queue.front());
        queue.pop_front();
      }
    }
    std::unique_lock<std::mutex> gl(gpu_mu);
    iter++;
    // ---- cancellation (ctx.Err() check per step, :156-159) ----
    for (size_t i = 0; i < active.size();) {
      if (active[i]->cancel) {
        release_stream_state(*active[i]);
        seq_free(active[i]->seq);
        delete reinterpret_cast<ptts_seq *>(active[i]->seq);
        active[i]->seq = nullptr;
        complete_request(active[i], PTTS_ECANCELED, "context canceled");
        active.erase(active.begin() + i);
      } else
        i++;
    }
    // ---- admission + batched prefill ----
    std::vector<std::shared_ptr<Request>> admitted, leftover;
    int ptoks = 0;
    for (auto &r : candidates) {
      if (r->cancel) {
        complete_request(r, PTTS_ECANCELED, "context canceled");
        continue;
      }
      if (!leftover.empty()) { // keep FIFO order once one had to wait
        leftover.push_back(r);
        continue;
      }
      int voice_tokens = r->voice_kv ? r->voice_kv->layers[0].offset : r->voice_frames;
      int prompt = (r->voice_kv ? 0 : r->voice_frames) + (int)r->tokens.size();
      int need = voice_tokens + (int)r->tokens.size() + r->max_steps;
      if (ptoks + prompt > max_rows) {
        leftover.push_back(r);
        continue;
      }
      Seq *s = nullptr;
      // shared voice prefix: map the cached pages, or (first use) become their donor; a
      // request whose voice is being prefilled by an earlier one of this batch waits a turn
      VoiceEntry *ve = r->voice_kv ? nullptr : voice_lookup(r->voice_emb, r->voice_frames, true);
      if (ve && !ve->ready && ve->refs < 0) {
        leftover.push_back(r);
        continue;
      }
      int rc;
      if (ve)
        ve->refs++; // pin: the evictions below must not take this entry
      if (ve && ve->ready) {
        rc = seq_alloc_shared(need, ve, &s);
        while (rc == PTTS_ENOMEM && voice_evict_one())
          rc = seq_alloc_shared(need, ve, &s);
        ve->refs--;
        if (rc == PTTS_OK)
          prompt -= s->kv_len; // only the unshared voice rows and the text are prefilled
      } else {
        rc = seq_alloc(need, &s);
        while (rc == PTTS_ENOMEM && voice_evict_one())
          rc = seq_alloc(need, &s);
        if (ve) {
          ve->refs--;
          if (rc == PTTS_OK)
            ve->refs = -1; // donor chosen: marks "prefill in flight" until voice_donate()
        }
      }
      r->voice_entry = ve;
      if (rc == PTTS_ENOMEM && (!active.empty() || !pending.empty() || !admitted.empty()) &&
          need <= cfg.max_seq_tokens) {
        leftover.push_back(r); // back-pressure: wait for pages / slots
        continue;
      }
      if (rc) {
        complete_request(r, rc, last_error());
        continue;
      }
      r->seq = s;
      ptoks += prompt;
      admitted.push_back(r);
    }
    if (!leftover.empty()) {
      std::unique_lock<std::mutex> lk(q_mu);
      for (auto it = leftover.rbegin(); it != leftover.rend(); ++it)
        queue.push_front(*it);
    }
    if (!admitted.empty()) {
      std::vector<Seq *> seqs;
      std::vector<const int64_t *> tok;
      std::vector<int> ntok, nvoice;
      std::vector<const float *> voice;
      int rc = PTTS_OK;
      // a request whose own inputs are bad fails alone (the reference fails only the offending
      // call); the rest of the batch is prefilled as usual
      for (size_t i = 0; i < admitted.size();) {
        auto &r = admitted[i];
        int vrc = r->voice_kv ? import_voice_kv(r->seq, r->voice_kv) : PTTS_OK;
        if (vrc) {
          std::string msg = last_error();
          if (r->voice_entry && !r->voice_entry->ready)
            r->voice_entry->refs = 0;
          seq_free(r->seq);
          delete reinterpret_cast<ptts_seq *>(r->seq);
          r->seq = nullptr;
          complete_request(r, vrc, msg.c_str());
          admitted.erase(admitted.begin() + i);
        } else
          i++;
      }
      for (auto &r : admitted) {
        auto now = std::chrono::steady_clock::now();
        r->res.queue_ms = std::chrono::duration<double, std::milli>(now - r->t_submit).count();
        cudaMemcpyAsync(d_noise + (size_t)r->seq->slot * max_frames * ldim, r->noise,
                        (size_t)r->max_steps * ldim * 4, cudaMemcpyHostToDevice, st);
        seqs.push_back(r->seq);
        tok.push_back(r->tokens.data());
        ntok.push_back((int)r->tokens.size());
        // a sequence on shared pages starts at kv_len = shared positions: skip those rows
        const int skip = r->seq->voice ? r->seq->kv_len : 0;
        voice.push_back(r->voice_kv || !r->voice_emb ? nullptr : r->voice_emb + (size_t)skip * D);
        nvoice.push_back(r->voice_kv ? 0 : r->voice_frames - skip);
      }
      if (!admitted.empty())
        rc = prefill(seqs, tok, ntok, voice, nvoice);
      for (auto &r : admitted)
        if (r->voice_entry && !r->voice_entry->ready) {
          r->voice_entry->refs = 0;
          if (rc == PTTS_OK)
            voice_donate(r->seq, r->voice_entry);
        }
      if (rc) {
        std::string msg = last_error();
        for (auto &r : admitted) {
          seq_free(r->seq);
          delete reinterpret_cast<ptts_seq *>(r->seq);
          r->seq = nullptr;
          complete_request(r, rc, msg.c_str());
        }
      } else {
        for (auto &r : admitted)
          active.push_back(r);
      }
    }
    // ---- one AR frame for every live sequence ----
    if (!active.empty()) {
      // group by lsd_steps (normally a single group)
      std::map<int, std::vector<size_t>> groups;
      for (size_t i = 0; i < active.size(); i++)
        groups[active[i]->lsd_steps].push_back(i);
      std::vector<char> finished(active.size(), 0), failed(active.size(), 0);
      std::string fail_msg;
      int fail_rc = 0;
      for (auto &kv : groups) {
        std::vector<Seq *> seqs;
        for (size_t i : kv.second)
          seqs.push_back(active[i]->seq);
        std::vector<float> eos(seqs.size());
        int rc = step(seqs, nullptr, nullptr, kv.first, nullptr, eos.data());
        if (rc) {
          fail_rc = rc;
          fail_msg = last_error();
          for (size_t i : kv.second)
            failed[i] = 1;
          continue;
        }
        auto now = std::chrono::steady_clock::now();
        for (size_t j = 0; j < kv.second.size(); j++) {
          Request &r = *active[kv.second[j]];
          const int stepi = r.seq->frames - 1; // 0-based step just executed
          r.eos.push_back(eos[j]);
          if (stepi == 0)
            r.res.first_frame_ms = std::chrono::duration<double, std::milli>(now - r.t_submit).count();
          // runtime_native_safetensors.go:178-190
          bool is_eos = eos[j] > (float)r.cfg.eos_threshold;
          if (r.cfg.force_eos_step >= 0 && stepi >= r.cfg.force_eos_step)
            is_eos = true;
          bool stop = false;
          if (is_eos && !r.have_cd) {
            r.have_cd = true;
            r.countdown = r.cfg.frames_after_eos;
          }
          if (r.have_cd) {
            if (r.countdown == 0)
              stop = true;
            else
              r.countdown--;
          }
          if (!stop && r.seq->frames >= r.max_steps)
            stop = true;
          if (stop)
            finished[kv.second[j]] = 1;
        }
      }
      std::vector<std::shared_ptr<Request>> keep;
      std::vector<DecodeItem> deliveries; // PTTS_REQ_STREAM: new frames + their receptive field
      const int wctx = mimi_window_frames();
      auto deliver = [&](const std::shared_ptr<Request> &r, bool final) {
        DecodeItem it;
        it.r = r;
        const int from = r->stream_issued, to = r->seq->frames;
        // Stateful delivery: the transformer's K/V of every earlier row sits in the request's ring, so only the
        // new frames plus ONE frame of conv left context (16 rows >= the 8 rows the conv stack reaches back) go
        // through the decoder.  Without a ring (none free, or a delivery too long for it): the bounded window.
        if (r->stream_state < 0 && from == 0 && !free_stream_states.empty()) {
          r->stream_state = free_stream_states.back();
          free_stream_states.pop_back();
        }
        const int sctx = from > 0 ? 1 : 0;
        if (r->stream_state >= 0 && (to - from + sctx) * up_stride <= MIMI_RING - Mctx && up_k <= 2 * up_stride &&
            Mctx > 0 && to * up_stride <= MIMI_ROPE_ROWS) {
          it.first = from - sctx;
          it.ring = r->stream_state;
          it.pre = it.first > 0 ? 1 : 0; // the context frame's rows take taps from the frame before it
        } else {
          if (r->stream_state >= 0) { // the ring has a gap from here on: give it back
            free_stream_states.push_back(r->stream_state);
            r->stream_state = -1;
          }
          it.first = from > wctx ? from - wctx : 0;
        }
        it.skip = from - it.first;
        it.count = to - it.first;
        it.final = final;
        r->stream_issued = to;
        if (final && r->stream_state >= 0) { // every later pass is ordered after this one on the Mimi stream
          free_stream_states.push_back(r->stream_state);
          r->stream_state = -1;
        }
        deliveries.push_back(it);
      };
      for (size_t i = 0; i < active.size(); i++) {
        auto &r = active[i];
        if (failed[i]) {
          release_stream_state(*r);
          seq_free(r->seq);
          delete reinterpret_cast<ptts_seq *>(r->seq);
          r->seq = nullptr;
          complete_request(r, fail_rc, fail_msg.c_str());
        } else if (finished[i]) {
          if (r->stream_frames > 0) { // the tail of a streaming request goes out at once
            r->res.n_frames = r->seq->frames;
            r->res.n_samples = (int64_t)r->seq->frames * info.samples_per_frame;
            deliver(r, true);
          } else {
            r->finished_at_iter = (int)iter;
            pending_frames += r->seq->frames;
            pending.push_back(r);
          }
        } else {
          if (r->stream_frames > 0 && r->seq->frames - r->stream_issued >= r->stream_frames)
            deliver(r, false);
          keep.push_back(r);
        }
      }
      if (!deliveries.empty()) {
        decode_async(deliveries);
        for (auto &it : deliveries)
          if (it.final) { // every read of the sequence's state is ordered on the stream already
            seq_free(it.r->seq);
            delete reinterpret_cast<ptts_seq *>(it.r->seq);
            it.r->seq = nullptr;
          }
      }
      active.swap(keep);
    }
    // ---- batched Mimi decode of finished chunks ----
    bool queue_empty;
    {
      std::unique_lock<std::mutex> lk(q_mu);
      queue_empty = queue.empty();
    }
    bool trigger = !pending.empty() &&
                   (pending_frames >= mimi_frames_cap / 2 || (active.empty() && queue_empty) ||
                    iter - pending.front()->finished_at_iter >= 16 ||
                    (int)(active.size() + pending.size()) >= max_batch);
    if (trigger) {
      for (auto &r : pending) {
        r->res.n_frames = r->seq->frames;
        r->res.n_samples = (int64_t)r->seq->frames * info.samples_per_frame;
      }
      std::vector<DecodeItem> whole;
      for (auto &r : pending) {
        DecodeItem it;
        it.r = r;
        it.count = r->seq->frames;
        whole.push_back(it);
      }
      decode_async(whole); // every request is completed by the finisher (or, on error, inside)
      for (auto &r : pending) {
        // every read of this sequence's state is ordered on the stream already
        seq_free(r->seq);
        delete reinterpret_cast<ptts_seq *>(r->seq);
        r->seq = nullptr;
      }
      pending.clear();
      pending_frames = 0;
    }
  }
}

} // namespace ptts

// ===========================================================================
// extern "C"
// ===========================================================================
using namespace ptts;

extern "C" {

const char *ptts_last_error(void) { return ptts::last_error(); }
const char *ptts_version(void) { return "pockettts-b200 native-cuda 0.1 (sm_100a)"; }

int ptts_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess)
    return 0;
  return n;
}

int ptts_create(const ptts_config *cfg, ptts_engine **out) {
  if (!out)
    PTTS_FAIL(PTTS_EINVAL, "native-cuda: ptts_create: out is NULL");
  *out = nullptr;
  ptts_config c{};
  if (cfg)
    c = *cfg;
  ptts_engine *e = new ptts_engine();
  int rc = e->impl.create(&c);
  if (rc) {
    delete e;
    return rc;
  }
  *out = e;
  return PTTS_OK;
}

void ptts_destroy(ptts_engine *e) {
  if (e)
    delete e;
}

int ptts_load_tensor(ptts_engine *e, const char *key, const int64_t *shape, int32_t rank,
                     int32_t dtype, const void *host_ptr) {
  if (!e)
    PTTS_FAIL(PTTS_EINVAL, "native-cuda runtime unavailable");
  return e->impl.load_tensor(key, shape, rank, dtype, host_ptr);
}

int ptts_load_safetensors(ptts_engine *e, const char *path) {
  if (!e)
    PTTS_FAIL(PTTS_EINVAL, "native-cuda runtime unavailable");
  return load_safetensors_file(&e->impl, path);
}

int ptts_finalize_weights(ptts_engine *e) {
  if (!e)
    PTTS_FAIL(PTTS_EINVAL, "native-cuda runtime unavailable");
  std::unique_lock<std::mutex> gl(e->impl.gpu_mu);
  int rc = e->impl.finalize();
  if (rc == PTTS_OK && cudaDeviceSynchronize() != cudaSuccess)
    PTTS_FAIL(PTTS_ECUDA, "cuda: device synchronize after weight upload failed");
  return rc;
}

int ptts_model_info_get(ptts_engine *e, ptts_model_info *out) {
  if (!e || !e->impl.finalized)
    PTTS_FAIL(PTTS_ESTATE, "native-cuda runtime unavailable");
  *out = e->impl.info;
  return PTTS_OK;
}

// flow_transformer.go:451-480,517-631: shape / offset guards of an upstream voice model state,
// checked at the call that carries it so a bad file fails only its own request.
static int validate_voice_kv(Engine &E, const ptts_voice_kv *vk) {
  if (vk->n_layers <= 0 || !vk->layers)
    PTTS_FAIL(PTTS_EINVAL, "generate: load voice model state: no layers");
  if (vk->n_layers != E.L)
    PTTS_FAIL(PTTS_EINVAL, "native: voice model state has %d layers, transformer has %d", vk->n_layers, E.L);
  for (int l = 0; l < vk->n_layers; l++) {
    const ptts_voice_layer &vl = vk->layers[l];
    if (!vl.cache)
      PTTS_FAIL(PTTS_EINVAL, "native: voice model state module \"transformer.layers.%d.self_attn\" missing cache", l);
    if (vl.offset < 0)
      PTTS_FAIL(PTTS_EINVAL, "native: voice model state module has negative offset %d", vl.offset);
    if (vl.offset > vl.cache_tokens)
      PTTS_FAIL(PTTS_EINVAL, "native: voice model state offset %d exceeds cache length %d", vl.offset,
                vl.cache_tokens);
    if (vl.offset != vk->layers[0].offset)
      PTTS_FAIL(PTTS_EINVAL, "native-cuda: voice model state offsets differ across layers");
  }
  if (vk->layers[0].offset > E.max_rows)
    PTTS_FAIL(PTTS_ENOMEM, "native-cuda: voice state (%d tokens) exceeds workspace rows", vk->layers[0].offset);
  return PTTS_OK;
}

static int validate_request(Engine &E, const int64_t *tokens, int32_t n_tokens, const float *voice_emb, int32_t voice_frames,
                            const ptts_voice_kv *voice_kv, const ptts_gen_cfg *cfg,
                            const float *noise, int32_t noise_rows, float *pcm_out, int64_t pcm_cap,
                            int *max_steps_out) {
  if (!E.finalized)
    PTTS_FAIL(PTTS_ESTATE, "native-cuda runtime unavailable");
  if (n_tokens <= 0)
    PTTS_FAIL(PTTS_EINVAL, "generate: token slice must not be empty");
  if (!cfg)
    PTTS_FAIL(PTTS_EINVAL, "generate: config is NULL");
  if (voice_emb && voice_kv)
    PTTS_FAIL(PTTS_EINVAL, "generate: voice embedding and voice model state are mutually exclusive");
  if (voice_emb && voice_frames <= 0)
    PTTS_FAIL(PTTS_EINVAL, "generate: build voice tensor: voice_frames must be > 0");
  if (voice_kv) {
    int rc = validate_voice_kv(E, voice_kv);
    if (rc)
      return rc;
  }
  if (!tokens)
    PTTS_FAIL(PTTS_EINVAL, "generate: NULL argument");
  // conditioner.go:31-53 range check, at the call (not inside the batched prefill)
  for (int j = 0; j < n_tokens; j++)
    if (tokens[j] < 0 || tokens[j] >= E.nbins)
      PTTS_FAIL(PTTS_EINVAL, "generate: text embeddings: native: token id %d (%lld) out of range [0,%d)", j,
                (long long)tokens[j], E.nbins);
  int ms = resolve_max_steps(*cfg, n_tokens);
  if (ms > E.max_frames)
    PTTS_FAIL(PTTS_ENOMEM, "generate: max steps %d exceeds engine max_frames %d", ms, E.max_frames);
  if (!noise || noise_rows < ms)
    PTTS_FAIL(PTTS_EINVAL, "generate: need %d noise rows, got %d", ms, noise_rows);
  bool no_pcm = (cfg->flags & PTTS_REQ_NO_PCM_D2H) != 0;
  if (!no_pcm && (!pcm_out || pcm_cap < (int64_t)ms * E.info.samples_per_frame))
    PTTS_FAIL(PTTS_EINVAL, "generate: pcm buffer too small (%lld < %lld)", (long long)pcm_cap,
              (long long)ms * E.info.samples_per_frame);
  *max_steps_out = ms;
  return PTTS_OK;
}

int ptts_submit(ptts_engine *e, const int64_t *tokens, int32_t n_tokens, const float *voice_emb,
                int32_t voice_frames, const ptts_voice_kv *voice_kv, const ptts_gen_cfg *cfg,
                const float *noise, int32_t noise_rows, float *pcm_out, int64_t pcm_cap,
                int64_t *ticket) {
  if (!e)
    PTTS_FAIL(PTTS_EINVAL, "native-cuda runtime unavailable");
  Engine &E = e->impl;
  int ms = 0;
  int rc = validate_request(E, tokens, n_tokens, voice_emb, voice_frames, voice_kv, cfg, noise, noise_rows,
                            pcm_out, pcm_cap, &ms);
  if (rc)
    return rc;
  if (!ticket)
    PTTS_FAIL(PTTS_EINVAL, "generate: NULL argument");
  auto r = std::make_shared<Request>();
  r->tokens.assign(tokens, tokens + n_tokens);
  r->voice_emb = voice_emb;
  r->voice_frames = voice_emb ? voice_frames : 0;
  r->voice_kv = voice_kv;
  r->cfg = *cfg;
  r->noise = noise;
  r->noise_rows = noise_rows;
  r->pcm_out = pcm_out;
  r->pcm_cap = pcm_cap;
  r->max_steps = ms;
  r->lsd_steps = cfg->lsd_decode_steps > 0 ? cfg->lsd_decode_steps : 1;
  if ((cfg->flags & PTTS_REQ_STREAM) && !(cfg->flags & PTTS_REQ_NO_PCM_D2H))
    r->stream_frames = cfg->stream_frames > 0 ? cfg->stream_frames : 1;
  r->t_submit = std::chrono::steady_clock::now();
  r->eng = &E;
  r->cost = (long)ms + n_tokens + (voice_emb ? voice_frames : 0);
  E.ensure_sched();
  {
    std::unique_lock<std::mutex> lk(E.q_mu);
    if (E.sched_stop)
      PTTS_FAIL(PTTS_ESTATE, "native-cuda runtime unavailable: engine is shutting down");
    E.inflight_cost.fetch_add(r->cost);
    E.inflight_requests.fetch_add(1);
    r->ticket = E.next_ticket++;
    E.tickets[r->ticket] = r;
    E.queue.push_back(r);
    *ticket = r->ticket;
  }
  E.q_cv.notify_all();
  return PTTS_OK;
}

static std::shared_ptr<Request> find_ticket(Engine &E, int64_t t) {
  std::unique_lock<std::mutex> lk(E.q_mu);
  auto it = E.tickets.find(t);
  return it == E.tickets.end() ? nullptr : it->second;
}

int ptts_wait(ptts_engine *e, int64_t ticket, ptts_result *result) {
  if (!e)
    PTTS_FAIL(PTTS_EINVAL, "native-cuda runtime unavailable");
  auto r = find_ticket(e->impl, ticket);
  if (!r)
    PTTS_FAIL(PTTS_EINVAL, "native-cuda: unknown ticket %lld", (long long)ticket);
  std::unique_lock<std::mutex> lk(r->mu);
  r->cv.wait(lk, [&] { return r->done; });
  if (result)
    *result = r->res;
  if (r->res.status != PTTS_OK)
    set_error("%s", r->err.c_str());
  return r->res.status;
}

int ptts_cancel(ptts_engine *e, int64_t ticket) {
  if (!e)
    PTTS_FAIL(PTTS_EINVAL, "native-cuda runtime unavailable");
  auto r = find_ticket(e->impl, ticket);
  if (!r)
    PTTS_FAIL(PTTS_EINVAL, "native-cuda: unknown ticket %lld", (long long)ticket);
  r->cancel = true;
  e->impl.q_cv.notify_all();
  return PTTS_OK;
}

int ptts_ticket_latents(ptts_engine *e, int64_t ticket, float *latents, float *eos_logits,
                        int32_t cap_frames) {
  if (!e)
    PTTS_FAIL(PTTS_EINVAL, "native-cuda runtime unavailable");
  auto r = find_ticket(e->impl, ticket);
  if (!r || !r->done)
    PTTS_FAIL(PTTS_ESTATE, "native-cuda: ticket not complete");
  int n = r->res.n_frames;
  if (cap_frames < n)
    PTTS_FAIL(PTTS_EINVAL, "native-cuda: latents buffer too small (%d < %d)", cap_frames, n);
  if (latents)
    memcpy(latents, r->latents.data(), (size_t)n * PTTS_LATENT_DIM * 4);
  if (eos_logits)
    memcpy(eos_logits, r->eos.data(), (size_t)n * 4);
  return n;
}

int ptts_ticket_error(ptts_engine *e, int64_t ticket, char *buf, int32_t cap) {
  if (!e || !buf || cap <= 0)
    PTTS_FAIL(PTTS_EINVAL, "native-cuda: ticket_error: bad arguments");
  auto r = find_ticket(e->impl, ticket);
  if (!r)
    PTTS_FAIL(PTTS_EINVAL, "native-cuda: unknown ticket %lld", (long long)ticket);
  std::unique_lock<std::mutex> lk(r->mu);
  const size_t n = r->err.size() < (size_t)cap - 1 ? r->err.size() : (size_t)cap - 1;
  memcpy(buf, r->err.data(), n);
  buf[n] = 0;
  return (int)n;
}

int ptts_ticket_read(ptts_engine *e, int64_t ticket, int32_t have_frames, int32_t timeout_ms, int32_t *frames_ready,
                     int32_t *done) {
  if (!e)
    PTTS_FAIL(PTTS_EINVAL, "native-cuda runtime unavailable");
  auto r = find_ticket(e->impl, ticket);
  if (!r)
    PTTS_FAIL(PTTS_EINVAL, "native-cuda: unknown ticket %lld", (long long)ticket);
  std::unique_lock<std::mutex> lk(r->mu);
  auto ready = [&] { return r->done || r->frames_ready > have_frames; };
  if (timeout_ms < 0)
    r->cv.wait(lk, ready);
  else
    r->cv.wait_for(lk, std::chrono::milliseconds(timeout_ms), ready);
  if (frames_ready)
    *frames_ready = r->done && r->res.status == PTTS_OK && r->stream_frames == 0 ? r->res.n_frames : r->frames_ready;
  if (done)
    *done = r->done ? 1 : 0;
  if (r->done && r->res.status != PTTS_OK) {
    set_error("%s", r->err.c_str());
    return r->res.status;
  }
  return PTTS_OK;
}

int ptts_ticket_release(ptts_engine *e, int64_t ticket) {
  if (!e)
    return PTTS_OK;
  std::unique_lock<std::mutex> lk(e->impl.q_mu);
  e->impl.tickets.erase(ticket);
  return PTTS_OK;
}

int ptts_generate_audio(ptts_engine *e, const int64_t *tokens, int32_t n_tokens,
                        const float *voice_emb, int32_t voice_frames, const ptts_voice_kv *voice_kv,
                        const ptts_gen_cfg *cfg, const float *noise, int32_t noise_rows,
                        float *pcm_out, int64_t pcm_cap, ptts_result *result) {
  int64_t t = 0;
  int rc = ptts_submit(e, tokens, n_tokens, voice_emb, voice_frames, voice_kv, cfg, noise, noise_rows,
                       pcm_out, pcm_cap, &t);
  if (rc)
    return rc;
  rc = ptts_wait(e, t, result);
  ptts_ticket_release(e, t);
  return rc;
}

// ---- stage-level API ------------------------------------------------------
int ptts_seq_open(ptts_engine *e, const int64_t *tokens, int32_t n_tokens, const float *voice_emb,
                  int32_t voice_frames, const ptts_voice_kv *voice_kv, int32_t reserve_frames,
                  ptts_seq **out) {
  if (!e || !e->impl.finalized)
    PTTS_FAIL(PTTS_ESTATE, "native-cuda runtime unavailable");
  if (!out)
    PTTS_FAIL(PTTS_EINVAL, "native-cuda: seq_open: out is NULL");
  if (voice_emb && voice_kv)
    PTTS_FAIL(PTTS_EINVAL, "generate: voice embedding and voice model state are mutually exclusive");
  Engine &E = e->impl;
  std::unique_lock<std::mutex> gl(E.gpu_mu);
  cudaSetDevice(E.device);
  if (reserve_frames <= 0)
    reserve_frames = E.max_frames;
  int vt = voice_kv ? voice_kv->layers[0].offset : (voice_emb ? voice_frames : 0);
  Seq *s = nullptr;
  int rc = E.seq_alloc(vt + n_tokens + reserve_frames, &s);
  if (rc)
    return rc;
  if (voice_kv)
    rc = E.import_voice_kv(s, voice_kv);
  if (rc == PTTS_OK && (n_tokens > 0 || (voice_emb && voice_frames > 0)))
    rc = E.prefill({s}, {tokens}, {n_tokens}, {voice_kv ? nullptr : voice_emb},
                   {voice_kv || !voice_emb ? 0 : voice_frames});
  if (rc) {
    E.seq_free(s);
    delete reinterpret_cast<ptts_seq *>(s);
    return rc;
  }
  *out = reinterpret_cast<ptts_seq *>(s);
  return PTTS_OK;
}

int ptts_step(ptts_engine *e, ptts_seq *const *seqs, int32_t batch, const float *frames_in,
              const float *noise, int32_t lsd_steps, float *latents_out, float *eos_logits_out) {
  if (!e || !e->impl.finalized)
    PTTS_FAIL(PTTS_ESTATE, "native-cuda runtime unavailable");
  if (!noise)
    PTTS_FAIL(PTTS_EINVAL, "native-cuda: step: noise is NULL");
  Engine &E = e->impl;
  std::unique_lock<std::mutex> gl(E.gpu_mu);
  cudaSetDevice(E.device);
  std::vector<Seq *> v;
  for (int i = 0; i < batch; i++) {
    if (!seqs[i])
      PTTS_FAIL(PTTS_EINVAL, "native: flow_lm state unavailable");
    v.push_back(&seqs[i]->impl);
  }
  return E.step(v, frames_in, noise, lsd_steps > 0 ? lsd_steps : 1, latents_out, eos_logits_out);
}

int ptts_decode(ptts_engine *e, ptts_seq *seq, int32_t first_frame, int32_t n_frames, float *pcm_out) {
  if (!e || !e->impl.finalized)
    PTTS_FAIL(PTTS_ESTATE, "native-cuda runtime unavailable");
  if (!seq || !pcm_out || first_frame < 0 || n_frames <= 0)
    PTTS_FAIL(PTTS_EINVAL, "native-cuda: decode: bad arguments");
  Engine &E = e->impl;
  std::unique_lock<std::mutex> gl(E.gpu_mu);
  cudaSetDevice(E.device);
  const int spf = E.info.samples_per_frame;
  if (first_frame == 0)
    return E.decode_seqs({&seq->impl}, {0}, {n_frames}, {pcm_out}, true);
  // Frame-level streaming in bounded work per emitted frame: the decoder is
  // causal with a finite receptive field — ML transformer layers of window Mctx rows (RoPE
  // is relative, LayerNorm per row), then convolutions reaching back a few rows — so the
  // samples of frame f depend only on the latent frames f - ctx .. f.  Decode that window
  // and hand back the requested tail.
  const int ctx = (E.ML * (E.Mctx > 0 ? E.Mctx - 1 : 8192) + 8 + E.up_stride - 1) / E.up_stride + 1;
  const int start = first_frame > ctx ? first_frame - ctx : 0;
  const int span = first_frame + n_frames - start;
  std::vector<float> tmp((size_t)span * spf);
  int rc = E.decode_seqs({&seq->impl}, {start}, {span}, {tmp.data()}, true);
  if (rc)
    return rc;
  memcpy(pcm_out, tmp.data() + (size_t)(first_frame - start) * spf, (size_t)n_frames * spf * 4);
  return PTTS_OK;
}

// mimi.go:14,791-794: the reference's native Mimi encoder is a hook that returns this error.
int ptts_mimi_encode_to_latent(ptts_engine *e, const float *, int64_t, float *, int32_t) {
  if (!e)
    PTTS_FAIL(PTTS_EINVAL, "native-cuda runtime unavailable");
  PTTS_FAIL(PTTS_ESTATE, "native mimi encoder is not implemented");
}

int ptts_decode_latents(ptts_engine *e, const float *latents, const int32_t *frames_per_seq,
                        int32_t n_seqs, float *pcm_out) {
  if (!e || !e->impl.finalized)
    PTTS_FAIL(PTTS_ESTATE, "native-cuda runtime unavailable");
  if (!latents || !frames_per_seq || n_seqs <= 0 || !pcm_out)
    PTTS_FAIL(PTTS_EINVAL, "native: latent tensor is nil");
  Engine &E = e->impl;
  std::unique_lock<std::mutex> gl(E.gpu_mu);
  cudaSetDevice(E.device);
  const int spf = E.info.samples_per_frame;
  const int max_segs = (E.mimi_rows_cap - E.up_stride * E.mimi_frames_cap) / E.mimi_pad;
  int i = 0;
  size_t foff = 0;
  while (i < n_seqs) {
    std::vector<MimiSeg> segs;
    int tf = 0;
    size_t f0 = foff;
    while (i < n_seqs && (int)segs.size() < max_segs) {
      int n = frames_per_seq[i];
      if (n <= 0)
        PTTS_FAIL(PTTS_EINVAL, "native: latent shape must be positive");
      if (n > E.mimi_frames_cap)
        PTTS_FAIL(PTTS_ENOMEM, "native-cuda: %d frames exceed mimi_group_frames %d", n, E.mimi_frames_cap);
      if (tf + n > E.mimi_frames_cap)
        break;
      MimiSeg sg;
      sg.n_frames = n;
      sg.frame0 = tf;
      sg.row0 = 0;
      sg.host_out = pcm_out + foff * spf;
      segs.push_back(sg);
      tf += n;
      foff += n;
      i++;
    }
    CUDA_TRY(cudaMemcpyAsync(E.m_latin, latents + f0 * E.ldim, (size_t)tf * E.ldim * 4,
                             cudaMemcpyHostToDevice, E.st));
    int rc = E.mimi_decode(segs, tf, true);
    if (rc)
      return rc;
  }
  return PTTS_OK;
}

int ptts_seq_frames(ptts_seq *s) { return s ? s->impl.frames : 0; }
int ptts_seq_kv_len(ptts_seq *s) { return s ? s->impl.kv_len : 0; }

int ptts_seq_read_kv(ptts_engine *e, ptts_seq *s, int32_t layer, int32_t which, float *out) {
  if (!e || !s || !out)
    PTTS_FAIL(PTTS_EINVAL, "native-cuda: read_kv: bad arguments");
  Engine &E = e->impl;
  std::unique_lock<std::mutex> gl(E.gpu_mu);
  cudaSetDevice(E.device);
  int T = s->impl.kv_len;
  if (T == 0)
    return PTTS_OK;
  if (layer < 0 || layer >= E.L || (size_t)T * E.D > (size_t)E.max_rows * E.D)
    PTTS_FAIL(PTTS_EINVAL, "native-cuda: read_kv: layer/size out of range");
  E.times.kernel_launches += kv_read_rows(E.nrm, T, s->impl.slot, E.kvg, layer, which ? 1 : 0, E.st);
  CUDA_TRY(cudaMemcpyAsync(out, E.nrm, (size_t)T * E.D * 4, cudaMemcpyDeviceToHost, E.st));
  CUDA_TRY(cudaStreamSynchronize(E.st));
  return PTTS_OK;
}

int ptts_seq_close(ptts_engine *e, ptts_seq *s) {
  if (!e || !s)
    return PTTS_OK;
  Engine &E = e->impl;
  std::unique_lock<std::mutex> gl(E.gpu_mu);
  E.seq_free(&s->impl);
  delete s;
  return PTTS_OK;
}

int ptts_stage_times_get(ptts_engine *e, ptts_stage_times *out, int32_t reset) {
  if (!e)
    PTTS_FAIL(PTTS_EINVAL, "native-cuda runtime unavailable");
  std::unique_lock<std::mutex> gl(e->impl.gpu_mu);
  if (out) {
    *out = e->impl.times;
    std::unique_lock<std::mutex> lk(e->impl.fin_mu);
    out->mimi_ms += e->impl.fin_mimi_ms;
    out->d2h_ms += e->impl.fin_d2h_ms;
  }
  if (reset) {
    e->impl.times = ptts_stage_times{};
    std::unique_lock<std::mutex> lk(e->impl.fin_mu);
    e->impl.fin_mimi_ms = 0;
    e->impl.fin_d2h_ms = 0;
  }
  return PTTS_OK;
}

int ptts_profile_enable(ptts_engine *e, int32_t on) {
  if (!e)
    PTTS_FAIL(PTTS_EINVAL, "native-cuda runtime unavailable");
  std::unique_lock<std::mutex> gl(e->impl.gpu_mu);
  e->impl.prof_on = on != 0;
  if (on)
    e->impl.prof_acc.clear();
  return PTTS_OK;
}

int ptts_profile_report(ptts_engine *e, char *buf, int32_t cap) {
  if (!e || !buf || cap <= 0)
    PTTS_FAIL(PTTS_EINVAL, "native-cuda: profile_report: bad arguments");
  std::unique_lock<std::mutex> gl(e->impl.gpu_mu);
  std::string s = "{";
  bool first = true;
  for (auto &kv : e->impl.prof_acc) {
    char tmp[512];
    snprintf(tmp, sizeof tmp, "%s\"%s\":{\"launches\":%lld,\"ms\":%.6f,\"flops\":%.6e,\"bytes\":%.6e}",
             first ? "" : ",", kv.first.c_str(), (long long)kv.second.launches, kv.second.ms,
             kv.second.flops, kv.second.bytes);
    s += tmp;
    first = false;
  }
  s += "}";
  if ((int)s.size() + 1 > cap)
    PTTS_FAIL(PTTS_EINVAL, "native-cuda: profile_report: buffer too small");
  memcpy(buf, s.c_str(), s.size() + 1);
  return (int)s.size();
}

int ptts_debug_gemm(ptts_engine *e, const float *A, int32_t M, int32_t K, const float *W, int32_t N,
                    const float *bias, int32_t act, int32_t path, float *C) {
  if (!e || !A || !W || !C || M <= 0 || N <= 0 || K <= 0)
    PTTS_FAIL(PTTS_EINVAL, "native-cuda: debug_gemm: bad arguments");
  std::unique_lock<std::mutex> gl(e->impl.gpu_mu);
  cudaSetDevice(e->impl.device);
  return e->impl.debug_gemm(A, M, K, W, N, bias, act, path, C);
}

int ptts_debug_gemm_tma(ptts_engine *e, const float *A, int32_t M, int32_t cin, int32_t taps, const float *W, int32_t N,
                        const float *bias, int32_t act, float *C, float *c_split) {
  if (!e || !A || !W || !C || M <= 0 || N <= 0 || cin <= 0 || taps <= 0)
    PTTS_FAIL(PTTS_EINVAL, "native-cuda: debug_gemm_tma: bad arguments");
  std::unique_lock<std::mutex> gl(e->impl.gpu_mu);
  cudaSetDevice(e->impl.device);
  return e->impl.debug_gemm_tma(A, M, cin, taps, W, N, bias, act, C, c_split);
}

int ptts_voice_cache_stats(ptts_engine *e, int64_t *hits, int64_t *shared_tokens, int32_t *entries) {
  if (!e)
    PTTS_FAIL(PTTS_EINVAL, "native-cuda runtime unavailable");
  std::unique_lock<std::mutex> gl(e->impl.gpu_mu);
  if (hits)
    *hits = e->impl.voice_hits;
  if (shared_tokens)
    *shared_tokens = e->impl.voice_shared_tokens;
  if (entries)
    *entries = (int32_t)e->impl.voice_cache.size();
  return PTTS_OK;
}

int ptts_debug_tc_force(int32_t bn, int32_t s) {
  ptts::tc_force_tile(bn, s);
  return PTTS_OK;
}

int ptts_debug_gemm_bench(ptts_engine *e, int32_t M, int32_t K, int32_t N, int32_t lda, int32_t path,
                          int32_t reps, float *avg_us) {
  if (!e || !avg_us || M <= 0 || N <= 0 || K <= 0 || reps <= 0)
    PTTS_FAIL(PTTS_EINVAL, "native-cuda: debug_gemm_bench: bad arguments");
  std::unique_lock<std::mutex> gl(e->impl.gpu_mu);
  cudaSetDevice(e->impl.device);
  return e->impl.debug_gemm_bench(M, K, N, lda, path, reps, avg_us);
}

int64_t ptts_debug_fetch(ptts_engine *e, const char *name, float *out, int64_t cap) {
  if (!e || !name)
    return -1;
  Engine &E = e->impl;
  std::unique_lock<std::mutex> gl(E.gpu_mu);
  cudaSetDevice(E.device);
  auto it = E.debug_bufs.find(name);
  if (it == E.debug_bufs.end())
    return -1;
  int64_t n = (int64_t)it->second.second;
  if (out) {
    int64_t c = n < cap ? n : cap;
    cudaMemcpyAsync(out, it->second.first, (size_t)c * 4, cudaMemcpyDeviceToHost, E.st);
    cudaStreamSynchronize(E.st);
  }
  return n;
}

} // extern "C"
