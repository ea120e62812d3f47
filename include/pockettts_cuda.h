/*
 * pockettts_cuda.h — C ABI of the B200-native `native-cuda` PocketTTS backend.
 *
 * This is the drop-in boundary for go-pocket-tts' synthesis hot path: the Go
 * side (internal/native/cuda, cgo) binds exactly these symbols and implements
 * `tts.Runtime` (reference internal/tts/runtime.go:42-45) on top of them.
 * Plain pointers and sizes only; no callbacks into the host language; every
 * pointer is caller-owned HOST memory that only has to stay valid for the call
 * (or, for ptts_submit, until the matching ptts_wait returns).
 *
 * Every entry point returns 0 on success or a negative PTTS_E* status;
 * ptts_last_error() returns the message for the calling thread.  The message
 * strings follow the reference's guard strings where a Go test asserts them
 * (internal/tts/runtime_native_safetensors_test.go:29-69).
 *
 * There is no CPU fallback: ptts_create fails when no sm_100 device is present.
 */
#ifndef POCKETTTS_CUDA_H
#define POCKETTTS_CUDA_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PTTS_OK 0
#define PTTS_EINVAL (-1)   /* bad argument / guard violation            */
#define PTTS_ECUDA (-2)    /* CUDA runtime error                        */
#define PTTS_ENOMEM (-3)   /* KV pages / slots / workspace exhausted    */
#define PTTS_ESTATE (-4)   /* call order violation (weights not final…) */
#define PTTS_ECANCELED (-5) /* request cancelled (ctx.Err() on Go side) */

#define PTTS_DTYPE_F32 0
#define PTTS_DTYPE_F16 1
#define PTTS_DTYPE_BF16 2

#define PTTS_LATENT_DIM 32
#define PTTS_SAMPLES_PER_FRAME 1920 /* 24000 Hz / 12.5 Hz, mimi.go:25-34 */

typedef struct ptts_engine ptts_engine;
typedef struct ptts_seq ptts_seq;

/* Engine construction.  Zero-initialise, then override; 0 means "default". */
typedef struct {
  int32_t device;          /* CUDA ordinal                                   */
  int32_t max_batch;       /* concurrently active sequences (default 256)    */
  int32_t page_tokens;     /* KV page size in tokens (default 16)            */
  int32_t max_seq_tokens;  /* KV tokens per sequence, <= 8192 (default 1024):
                              RoPE table cap, flow_transformer.go:505         */
  int32_t kv_pages;        /* KV pool pages (default max_batch*max_seq/page)  */
  int32_t max_frames;      /* latent frames per sequence (default 512)        */
  int32_t mimi_group_frames; /* Mimi decode workspace, in latent frames per
                              batched decode (default 2048)                   */
  int32_t max_prefill_tokens; /* prompt tokens per batched prefill pass
                              (default 8192)                                  */
  int32_t flags;           /* PTTS_ENGINE_* */
} ptts_config;

#define PTTS_ENGINE_NO_GRAPHS 1 /* disable CUDA-graph replay of the B<=8 step */
#define PTTS_ENGINE_NO_TENSOR_CORES 2 /* keep every GEMM on the fp32 CUDA-core path */
#define PTTS_ENGINE_KV_F32 4 /* FlowLM KV cache as fp32 like the reference (flow_transformer.go:26-108);
                                default: fp16 storage, fp32 attention arithmetic (half the cache stream) */

/* Per-call generation knobs: mirrors tts.RuntimeGenerateConfig
 * (reference internal/tts/runtime.go:11-31). */
typedef struct {
  double temperature;   /* informational; noise arrives already scaled by
                           sqrt(temperature) (flow_lm.go:386-408)             */
  double eos_threshold; /* EOS iff raw logit > threshold (flow_lm.go:281)     */
  int32_t max_steps;
  int32_t estimated_max_steps;
  int32_t lsd_decode_steps;
  int32_t frames_after_eos;
  int32_t force_eos_step; /* <0: off.  Synthetic-benchmark extension: declare
                           EOS at this 0-based step (random-init weights have
                           no meaningful EOS logit)            */
  int32_t flags;          /* PTTS_REQ_* */
  int32_t stream_frames;  /* PTTS_REQ_STREAM: emit PCM every this many latent frames (<= 0: 1) */
} ptts_gen_cfg;

#define PTTS_REQ_NO_PCM_D2H 1 /* leave PCM on the device (bench "value" leg) */
#define PTTS_REQ_PCM16 2      /* pcm_out is int16_t[pcm_cap]: clamp(v,-1,1)*32767 truncated on the device
                                 (internal/audio/wav_stream.go:44-54), half the D2H bytes */

#define PTTS_REQ_STREAM 4     /* frame-level streaming (reference PLAN.md:9-17, service.go:161-200,
                                 tts/runtime.go:33-38 PCMChunk): while the AR loop of the request is still running,
                                 every `stream_frames` new latent frames are decoded — only their receptive field,
                                 the previous 33 frames, is re-run through Mimi — and their samples land in pcm_out
                                 at their final position; ptts_ticket_read tells the caller how far the PCM is
                                 valid.  The concatenation equals the whole-chunk decode (causal decoder). */

/* Upstream voice model state for one transformer layer: the
 * `transformer.layers.{i}.self_attn/cache` tensor, layout [2,1,T,H,Dh] fp32
 * (K half then V half), and `/offset` (flow_transformer.go:517-631). */
typedef struct {
  const float *cache;
  int32_t cache_tokens; /* T */
  int32_t offset;       /* valid tokens, 0 <= offset <= T */
} ptts_voice_layer;

typedef struct {
  const ptts_voice_layer *layers; /* one per FlowLM layer */
  int32_t n_layers;
} ptts_voice_kv;

typedef struct {
  int32_t d_model, n_heads, head_dim, n_layers, ffn_dim;
  int32_t flow_dim, flow_depth, latent_dim, n_bins;
  int32_t mimi_dim, mimi_heads, mimi_layers, mimi_ffn_dim, mimi_context;
  int32_t samples_per_frame, mimi_steps_per_latent;
  int64_t params_ar;   /* weights streamed by one AR frame (SURVEY §8d P_AR)  */
  int64_t params_mimi; /* Mimi decoder + projector weights                    */
  int64_t macs_ar_frame;   /* multiply-accumulates per sequence-frame, AR     */
  int64_t macs_mimi_frame; /* ... Mimi decode (without window attention)      */
  int64_t macs_prefill_token;
} ptts_model_info;

/* CUDA-event stage times, accumulated since the last reset (milliseconds). */
typedef struct {
  double prefill_ms, step_ms, flow_head_ms, mimi_ms, d2h_ms;
  int64_t prefill_tokens, step_rows, steps, mimi_frames;
  int64_t kernel_launches; /* kernels this engine launched */
} ptts_stage_times;

/* Result of one request. */
typedef struct {
  int32_t status;   /* PTTS_OK or error */
  int32_t n_frames; /* latent frames generated */
  int64_t n_samples; /* 1920 * n_frames */
  double queue_ms;       /* submit -> admitted                            */
  double first_frame_ms; /* submit -> first latent frame done (host clock) */
  double total_ms;       /* submit -> PCM on host                          */
} ptts_result;

const char *ptts_last_error(void);
const char *ptts_version(void);
int ptts_device_count(void);

/* ---- engine + weights: replaces native.LoadModelFromSafetensors
 *      (reference internal/native/model.go:33-66) ---------------------------- */
int ptts_create(const ptts_config *cfg, ptts_engine **out);
void ptts_destroy(ptts_engine *e); /* NULL-safe, idempotent per handle */

/* One call per safetensors key, as the Go loader walks the store
 * (internal/safetensors/store.go:195-211, internal/native/var_builder.go:62-84).
 * F16/BF16 payloads are widened to fp32 exactly as store.go:339-395 does. */
int ptts_load_tensor(ptts_engine *e, const char *key, const int64_t *shape,
                     int32_t rank, int32_t dtype, const void *host_ptr);
/* Convenience for non-Go hosts: parse a .safetensors file and load every key. */
int ptts_load_safetensors(ptts_engine *e, const char *path);
/* Probes layer counts (flow_transformer.go:485-497, flow_net.go:273-285,
 * mimi.go:480-492), builds RoPE tables, folds the latent->Mimi projector
 * (model.go:176-250), repacks conv kernels, allocates KV pool + workspaces. */
int ptts_finalize_weights(ptts_engine *e);
int ptts_model_info_get(ptts_engine *e, ptts_model_info *out);

/* ---- tts.Runtime.GenerateAudio
 *      (reference internal/tts/runtime_native_safetensors.go:52-238) ---------
 * Blocking and thread-safe: concurrent callers are continuously batched by the
 * engine's scheduler thread (this replaces the rngMu-serialised AR step,
 * :161-170, and the server's semaphore pool).
 *   tokens/n_tokens : SentencePiece ids (borrowed)
 *   voice_emb       : [voice_frames, d_model] prefix or NULL (:104-119)
 *   voice_kv        : upstream model state or NULL (:123-136); exclusive with
 *                     voice_emb (:100-102)
 *   noise           : [noise_rows, 32] host-drawn N(0, temperature) rows, one
 *                     per AR step, noise_rows >= resolved max steps
 *   pcm_out/pcm_cap : caller buffer, >= 1920 * resolved max steps samples
 */
int ptts_generate_audio(ptts_engine *e, const int64_t *tokens, int32_t n_tokens,
                        const float *voice_emb, int32_t voice_frames,
                        const ptts_voice_kv *voice_kv, const ptts_gen_cfg *cfg,
                        const float *noise, int32_t noise_rows, float *pcm_out,
                        int64_t pcm_cap, ptts_result *result);

/* Asynchronous form of the same call (what a Go scheduler goroutine or the
 * benchmark uses to keep many requests in flight).  Buffers must stay valid
 * until ptts_wait returns for the ticket. */
int ptts_submit(ptts_engine *e, const int64_t *tokens, int32_t n_tokens,
                const float *voice_emb, int32_t voice_frames,
                const ptts_voice_kv *voice_kv, const ptts_gen_cfg *cfg,
                const float *noise, int32_t noise_rows, float *pcm_out,
                int64_t pcm_cap, int64_t *ticket);
int ptts_wait(ptts_engine *e, int64_t ticket, ptts_result *result);
/* ctx cancellation (runtime_native_safetensors.go:156-159): the sequence is
 * evicted at the next step boundary and ptts_wait returns PTTS_ECANCELED. */
int ptts_cancel(ptts_engine *e, int64_t ticket);
/* Optional latents tap for parity harnesses: copies the request's latent
 * frames [n_frames,32] and EOS logits [n_frames]; valid after ptts_wait and
 * until the ticket is released. */
int ptts_ticket_latents(ptts_engine *e, int64_t ticket, float *latents,
                        float *eos_logits, int32_t cap_frames);
/* The request's own error text (empty when it succeeded), copied into buf (NUL-terminated,
 * truncated to cap); returns its length.  Unlike ptts_last_error() it does not depend on
 * which OS thread asks: a Go caller reads it after ptts_wait without runtime.LockOSThread
 * (errors of GenerateAudio, runtime_native_safetensors.go:91,173,214). */
int ptts_ticket_error(ptts_engine *e, int64_t ticket, char *buf, int32_t cap);
/* Drains a PTTS_REQ_STREAM request: blocks until more than `have_frames` frames of PCM are valid in
 * the request's pcm_out (samples [0, 1920 * *frames_ready)), the request is complete, or timeout_ms
 * (< 0: no timeout) has passed.  *done is set once every frame has been delivered (or the request
 * failed: then the return value is its status and ptts_ticket_error has the text). */
int ptts_ticket_read(ptts_engine *e, int64_t ticket, int32_t have_frames, int32_t timeout_ms,
                     int32_t *frames_ready, int32_t *done);
int ptts_ticket_release(ptts_engine *e, int64_t ticket);

/* ---- multi-GPU pool: one process, one engine per device -------------------
 * Replaces the server's worker semaphore on a multi-GPU box (reference
 * internal/server/server.go:114,402-421): every device holds a full weight replica, its own
 * paged KV pool and scheduler thread; a submitted chunk goes to the least-loaded device (frames
 * still to generate + prompt tokens still to prefill over its unfinished requests).  Chunks are
 * independent (internal/tts/service.go:138-153), so nothing crosses GPUs: no collective.
 * devices == NULL / n_devices == 0: every visible GPU.  cfg->device is ignored. */
typedef struct ptts_pool ptts_pool;
int ptts_pool_create(const ptts_config *cfg, const int32_t *devices, int32_t n_devices, ptts_pool **out);
void ptts_pool_destroy(ptts_pool *p); /* NULL-safe */
int ptts_pool_size(ptts_pool *p);
/* Borrows engine `index` (model info, stage times, voice cache stats); owned by the pool. */
int ptts_pool_engine(ptts_pool *p, int32_t index, ptts_engine **out);
int ptts_pool_load_tensor(ptts_pool *p, const char *key, const int64_t *shape, int32_t rank, int32_t dtype,
                          const void *host_ptr);
int ptts_pool_load_safetensors(ptts_pool *p, const char *path);
int ptts_pool_finalize_weights(ptts_pool *p);
/* Same contract as ptts_submit / ptts_wait / ...; engine_index (optional) receives the device slot
 * the chunk was sent to.  A multi-chunk utterance may fan out over GPUs: the caller re-assembles
 * its chunks in order (service.go:138-153 concatenates them anyway). */
int ptts_pool_submit(ptts_pool *p, const int64_t *tokens, int32_t n_tokens, const float *voice_emb,
                     int32_t voice_frames, const ptts_voice_kv *voice_kv, const ptts_gen_cfg *cfg,
                     const float *noise, int32_t noise_rows, float *pcm_out, int64_t pcm_cap, int64_t *ticket,
                     int32_t *engine_index);
int ptts_pool_wait(ptts_pool *p, int64_t ticket, ptts_result *result);
int ptts_pool_cancel(ptts_pool *p, int64_t ticket);
int ptts_pool_ticket_error(ptts_pool *p, int64_t ticket, char *buf, int32_t cap);
int ptts_pool_ticket_latents(ptts_pool *p, int64_t ticket, float *latents, float *eos_logits, int32_t cap_frames);
int ptts_pool_ticket_read(ptts_pool *p, int64_t ticket, int32_t have_frames, int32_t timeout_ms,
                          int32_t *frames_ready, int32_t *done);
int ptts_pool_ticket_release(ptts_pool *p, int64_t ticket);
int ptts_pool_generate_audio(ptts_pool *p, const int64_t *tokens, int32_t n_tokens, const float *voice_emb,
                             int32_t voice_frames, const ptts_voice_kv *voice_kv, const ptts_gen_cfg *cfg,
                             const float *noise, int32_t noise_rows, float *pcm_out, int64_t pcm_cap,
                             ptts_result *result);
/* Dispatcher view, one entry per engine (up to cap): outstanding load, unfinished requests,
 * requests dispatched so far.  Returns the number of entries written. */
int ptts_pool_load(ptts_pool *p, int64_t *inflight_cost, int32_t *inflight_requests, int64_t *dispatched,
                   int32_t cap);

/* ---- stage-level API: mirrors native.Model
 *      {NewFlowState, NewFlowStateFromVoiceModelState, PromptFlow,
 *       SampleNextLatentStateful, LatentToMimi, MimiDecode}
 *      (reference internal/native/model.go:110-167,410-416) ------------------ */
/* Opens a sequence: allocates a slot + KV pages, imports voice KV if given,
 * gathers text embeddings (conditioner.go:31-53), prepends voice_emb and runs
 * the prefill (flow_transformer.go:749-771). */
int ptts_seq_open(ptts_engine *e, const int64_t *tokens, int32_t n_tokens,
                  const float *voice_emb, int32_t voice_frames,
                  const ptts_voice_kv *voice_kv, int32_t reserve_frames,
                  ptts_seq **out);
/* One AR frame for B sequences at once (flow_lm.go:238-299).
 *   frames_in : [B,32] previous latents, NaN lanes = BOS (tensor_util.go:242-271),
 *               or NULL to feed back each sequence's own last output
 *   noise     : [B,32]
 *   latents_out [B,32], eos_logits_out [B] (either may be NULL) */
int ptts_step(ptts_engine *e, ptts_seq *const *seqs, int32_t batch,
              const float *frames_in, const float *noise, int32_t lsd_steps,
              float *latents_out, float *eos_logits_out);
/* Decodes latent frames [first_frame, first_frame+n_frames) of the sequence's
 * own history to PCM (model.go:141-167 + mimi.go:719-789).  The Mimi stack is
 * causal, so a prefix decode equals the same span of the whole-chunk decode. */
int ptts_decode(ptts_engine *e, ptts_seq *seq, int32_t first_frame,
                int32_t n_frames, float *pcm_out);
/* Stateless decode of caller-supplied latents [n_frames,32] for several
 * independent chunks at once (ragged batch): frames_per_seq[n_seqs]. */
int ptts_decode_latents(ptts_engine *e, const float *latents,
                        const int32_t *frames_per_seq, int32_t n_seqs,
                        float *pcm_out);
/* MimiModel.EncodeToLatent (internal/native/mimi.go:791-794): the reference's native backend leaves the Mimi
 * encoder as a hook and returns ErrMimiEncoderNotImplemented ("native mimi encoder is not implemented",
 * mimi.go:14); voices come in as embeddings or model states (a9).  Mirrored with the same behaviour: always
 * PTTS_ESTATE and that message, nothing is written to latents_out. */
int ptts_mimi_encode_to_latent(ptts_engine *e, const float *pcm, int64_t n_samples, float *latents_out,
                               int32_t cap_frames);
int ptts_seq_frames(ptts_seq *s);
int ptts_seq_kv_len(ptts_seq *s);
/* Parity tap: valid K or V rows of one layer as [H, kv_len, Dh]. */
int ptts_seq_read_kv(ptts_engine *e, ptts_seq *s, int32_t layer, int32_t which,
                     float *out);
int ptts_seq_close(ptts_engine *e, ptts_seq *s);

/* ---- instrumentation (bench/stageprof "device-timed stages") -------------- */
int ptts_stage_times_get(ptts_engine *e, ptts_stage_times *out, int32_t reset);
/* Per-kernel-class CUDA-event profiling: while enabled every GEMM / attention
 * launch is bracketed by an event pair on the launch stream; the report is a
 * JSON object {class: {launches, ms, flops, bytes}} with ALGORITHMIC flops and
 * bytes. */
int ptts_profile_enable(ptts_engine *e, int32_t on);
int ptts_profile_report(ptts_engine *e, char *buf, int32_t cap);
/* Stand-alone C[M,N] = act(A[M,K] x W[N,K]^T + bias) through a chosen kernel
 * path (0 = auto, 1 = CUDA-core SGEMM/GEMV, 2 = tcgen05 split-bf16): kernel
 * unit-test hook. */
int ptts_debug_gemm(ptts_engine *e, const float *A, int32_t M, int32_t K, const float *W, int32_t N,
                    const float *bias, int32_t act, int32_t path, float *C);
/* Times one GEMM shape on device data: average microseconds over `reps`
 * launches (CUDA events on the engine stream).  lda < K = overlapping rows. */
/* Voice prefix cache: requests that carry the same `audio_prompt` embedding share the
 * KV pages of its first floor(T/16)*16 positions; hits = sequences admitted on shared pages,
 * shared_tokens = prefill positions they did not recompute, entries = voices currently cached. */
int ptts_voice_cache_stats(ptts_engine *e, int64_t *hits, int64_t *shared_tokens, int32_t *entries);
/* Unit-test hook of the TMA-fed GEMM: A [M,cin] is split into bf16 planes on the device;
 * taps > 1 = implicit-GEMM causal conv (W [N, taps*cin], row m sees rows m-taps+1..m, rows < 0 zero).
 * c_split (optional) receives the split-plane copy of the output (hi + lo). */
int ptts_debug_gemm_tma(ptts_engine *e, const float *A, int32_t M, int32_t cin, int32_t taps, const float *W,
                        int32_t N, const float *bias, int32_t act, float *C, float *c_split);
/* Sweep hook: force tile width / split-K of the tcgen05 GEMM (bn <= 0 restores the heuristic). */
int ptts_debug_tc_force(int32_t bn, int32_t s);
int ptts_debug_gemm_bench(ptts_engine *e, int32_t M, int32_t K, int32_t N, int32_t lda, int32_t path,
                          int32_t reps, float *avg_us);
/* Copies a named internal buffer of the last pass (debug / stage parity). */
int64_t ptts_debug_fetch(ptts_engine *e, const char *name, float *out,
                         int64_t cap);

#ifdef __cplusplus
}
#endif
#endif /* POCKETTTS_CUDA_H */
