//go:build cgo && linux

// Package cuda binds the B200-native PocketTTS engine (libpockettts_cuda.so)
// for the `native-cuda` backend.  It replaces internal/native.Model for the
// synthesis hot path: weights are uploaded once from the safetensors store,
// GenerateAudio becomes one blocking C call that the engine batches
// continuously with every other in-flight request.
package cuda

/*
#cgo CFLAGS: -I${SRCDIR}/../../../third_party/pockettts_b200/include
#cgo LDFLAGS: -L${SRCDIR}/../../../third_party/pockettts_b200/lib -lpockettts_cuda -Wl,-rpath,$ORIGIN/../lib
#include <stdlib.h>
#include "pockettts_cuda.h"
*/
import "C"

import (
	"errors"
	"fmt"
	"runtime"
	"unsafe"

	"github.com/cwbudde/go-pocket-tts/internal/safetensors"
)

// ErrCanceled is returned when the request's context was cancelled.
var ErrCanceled = errors.New("native-cuda: request canceled")

// Config mirrors ptts_config.
type Config struct {
	Device           int
	MaxBatch         int
	MaxSeqTokens     int
	MaxFrames        int
	KVPages          int
	MimiGroupFrames  int
	MaxPrefillTokens int
}

// GenConfig mirrors ptts_gen_cfg (tts.RuntimeGenerateConfig minus the voices).
type GenConfig struct {
	Temperature       float64
	EOSThreshold      float64
	MaxSteps          int
	EstimatedMaxSteps int
	LSDDecodeSteps    int
	FramesAfterEOS    int
	// StreamFrames > 0 asks for frame-level delivery (PTTS_REQ_STREAM): PCM becomes valid in
	// the ticket's buffer every StreamFrames latent frames; see Ticket.Read.
	StreamFrames int
}

// VoiceLayer is one `transformer.layers.{i}.self_attn` module of an upstream
// voice model-state file: cache [2,1,T,H,D] and offset.
type VoiceLayer struct {
	Cache       []float32
	CacheTokens int
	Offset      int
}

// Engine owns the device-resident model of one GPU.
type Engine struct {
	h *C.ptts_engine
}

// lastError reads the calling thread's message.  ptts_last_error() is thread-local on
// the C side, so every caller brackets the failing call and this read with
// runtime.LockOSThread (see locked); request errors do not use it at all — Wait reads the
// message stored on the ticket (ptts_ticket_error).
func lastError(rc C.int) error {
	if rc == C.PTTS_ECANCELED {
		return ErrCanceled
	}
	return errors.New(C.GoString(C.ptts_last_error()))
}

// locked runs one C call and, if it failed, fetches its message on the same OS thread.
func locked(call func() C.int) error {
	runtime.LockOSThread()
	defer runtime.UnlockOSThread()
	if rc := call(); rc != 0 {
		return lastError(rc)
	}
	return nil
}

// Open creates the engine and uploads every tensor of the store, exactly the
// keys internal/native loads (LoadModelFromStore, model.go:42-66).
func Open(store *safetensors.Store, cfg Config) (*Engine, error) {
	var cc C.ptts_config
	cc.device = C.int32_t(cfg.Device)
	cc.max_batch = C.int32_t(cfg.MaxBatch)
	cc.max_seq_tokens = C.int32_t(cfg.MaxSeqTokens)
	cc.max_frames = C.int32_t(cfg.MaxFrames)
	cc.kv_pages = C.int32_t(cfg.KVPages)
	cc.mimi_group_frames = C.int32_t(cfg.MimiGroupFrames)
	cc.max_prefill_tokens = C.int32_t(cfg.MaxPrefillTokens)

	e := &Engine{}
	if err := locked(func() C.int { return C.ptts_create(&cc, &e.h) }); err != nil {
		return nil, fmt.Errorf("native-cuda: create engine: %w", err)
	}

	for _, name := range store.Names() {
		t, err := store.Tensor(name) // decoded to []float32 (store.go:195-211)
		if err != nil {
			e.Close()
			return nil, err
		}
		if len(t.Data) == 0 {
			continue
		}
		cname := C.CString(name)
		shape := make([]C.int64_t, len(t.Shape))
		for i, d := range t.Shape {
			shape[i] = C.int64_t(d)
		}
		var sp *C.int64_t
		if len(shape) > 0 {
			sp = &shape[0]
		}
		err = locked(func() C.int {
			return C.ptts_load_tensor(e.h, cname, sp, C.int32_t(len(shape)), C.PTTS_DTYPE_F32,
				unsafe.Pointer(&t.Data[0]))
		})
		C.free(unsafe.Pointer(cname))
		runtime.KeepAlive(t)
		if err != nil {
			e.Close()
			return nil, fmt.Errorf("native-cuda: upload %q: %w", name, err)
		}
	}

	if err := locked(func() C.int { return C.ptts_finalize_weights(e.h) }); err != nil {
		e.Close()
		return nil, fmt.Errorf("native-cuda: finalize weights: %w", err)
	}
	return e, nil
}

// Close releases the engine. Nil-safe and idempotent.
func (e *Engine) Close() {
	if e == nil || e.h == nil {
		return
	}
	C.ptts_destroy(e.h)
	e.h = nil
}

// Ticket identifies an in-flight request.
type Ticket struct {
	id   C.int64_t
	pcm  []float32
	pins runtime.Pinner
	// C-allocated voice descriptor (the engine reads it until Wait returns)
	vkv    *C.ptts_voice_kv
	layers unsafe.Pointer
}

func (t *Ticket) release() {
	t.pins.Unpin()
	if t.layers != nil {
		C.free(t.layers)
		t.layers = nil
	}
	if t.vkv != nil {
		C.free(unsafe.Pointer(t.vkv))
		t.vkv = nil
	}
}

// Submit queues one chunk.  noise is [maxSteps*32], drawn by the caller
// (flow_lm.go:386-408); voiceEmb is [Tv*1024] or nil; voice is the upstream
// model state or nil.  All slices stay pinned until Wait returns.
func (e *Engine) Submit(tokens []int64, voiceEmb []float32, voiceFrames int, voice []VoiceLayer,
	cfg GenConfig, noise []float32, maxSteps int) (*Ticket, error) {
	if e == nil || e.h == nil {
		return nil, errors.New("native-cuda runtime unavailable")
	}
	if len(tokens) == 0 {
		return nil, errors.New("generate: token slice must not be empty")
	}

	t := &Ticket{pcm: make([]float32, maxSteps*C.PTTS_SAMPLES_PER_FRAME)}
	var g C.ptts_gen_cfg
	g.temperature = C.double(cfg.Temperature)
	g.eos_threshold = C.double(cfg.EOSThreshold)
	g.max_steps = C.int32_t(maxSteps)
	g.estimated_max_steps = C.int32_t(cfg.EstimatedMaxSteps)
	g.lsd_decode_steps = C.int32_t(cfg.LSDDecodeSteps)
	g.frames_after_eos = C.int32_t(cfg.FramesAfterEOS)
	g.force_eos_step = -1
	if cfg.StreamFrames > 0 {
		g.flags |= C.PTTS_REQ_STREAM
		g.stream_frames = C.int32_t(cfg.StreamFrames)
	}

	t.pins.Pin(&tokens[0])
	t.pins.Pin(&noise[0])
	t.pins.Pin(&t.pcm[0])

	var vemb *C.float
	if len(voiceEmb) > 0 {
		t.pins.Pin(&voiceEmb[0])
		vemb = (*C.float)(unsafe.Pointer(&voiceEmb[0]))
	}

	var vkv *C.ptts_voice_kv
	if len(voice) > 0 {
		// C-allocated so the engine may keep the descriptor until Wait.
		layers := (*[1 << 10]C.ptts_voice_layer)(C.calloc(C.size_t(len(voice)), C.sizeof_ptts_voice_layer))
		for i := range voice {
			t.pins.Pin(&voice[i].Cache[0])
			layers[i].cache = (*C.float)(unsafe.Pointer(&voice[i].Cache[0]))
			layers[i].cache_tokens = C.int32_t(voice[i].CacheTokens)
			layers[i].offset = C.int32_t(voice[i].Offset)
		}
		vkv = (*C.ptts_voice_kv)(C.calloc(1, C.sizeof_ptts_voice_kv))
		vkv.layers = &layers[0]
		vkv.n_layers = C.int32_t(len(voice))
		t.vkv, t.layers = vkv, unsafe.Pointer(layers)
	}

	err := locked(func() C.int {
		return C.ptts_submit(e.h, (*C.int64_t)(unsafe.Pointer(&tokens[0])), C.int32_t(len(tokens)),
			vemb, C.int32_t(voiceFrames), vkv, &g,
			(*C.float)(unsafe.Pointer(&noise[0])), C.int32_t(len(noise)/C.PTTS_LATENT_DIM),
			(*C.float)(unsafe.Pointer(&t.pcm[0])), C.int64_t(len(t.pcm)), &t.id)
	})
	if err != nil {
		t.release()
		return nil, err
	}
	return t, nil
}

// Wait blocks until the chunk is synthesised and returns a fresh PCM slice
// owned by the caller (runtime_native_safetensors.go:237).
func (e *Engine) Wait(t *Ticket) ([]float32, int, error) {
	var res C.ptts_result
	rc := C.ptts_wait(e.h, t.id, &res)
	var err error
	if rc == C.PTTS_ECANCELED {
		err = ErrCanceled
	} else if rc != 0 {
		// the message lives on the ticket, not in thread-local state: the goroutine may have
		// moved to another OS thread while it was blocked in ptts_wait
		var buf [512]C.char
		C.ptts_ticket_error(e.h, t.id, &buf[0], C.int32_t(len(buf)))
		err = errors.New(C.GoString(&buf[0]))
	}
	C.ptts_ticket_release(e.h, t.id)
	t.release()
	if err != nil {
		return nil, 0, err
	}
	out := append([]float32(nil), t.pcm[:int(res.n_samples)]...)
	return out, int(res.n_frames), nil
}

// Cancel evicts the request at the next step boundary.
func (e *Engine) Cancel(t *Ticket) { C.ptts_cancel(e.h, t.id) }

// EncodeToLatent mirrors native.MimiModel.EncodeToLatent (internal/native/mimi.go:791-794): the native backend
// has no Mimi encoder; the error text is ErrMimiEncoderNotImplemented's.
func (e *Engine) EncodeToLatent(pcm []float32) ([]float32, error) {
	var p *C.float
	if len(pcm) > 0 {
		p = (*C.float)(unsafe.Pointer(&pcm[0]))
	}
	err := locked(func() C.int {
		return C.ptts_mimi_encode_to_latent(e.h, p, C.int64_t(len(pcm)), nil, 0)
	})
	return nil, err
}

// Read blocks until more than have frames of PCM are valid (or the request is done) and returns the
// new frame count; samples [have*1920, ready*1920) of the returned slice are the freshly decoded
// audio.  This is what lets Service.SynthesizeStream (internal/tts/service.go:161-200) emit a
// PCMChunk (internal/tts/runtime.go:33-38) per frame instead of per text chunk.
func (e *Engine) Read(t *Ticket, have int) (pcm []float32, ready int, done bool, err error) {
	var r, d C.int32_t
	rc := C.ptts_ticket_read(e.h, t.id, C.int32_t(have), -1, &r, &d)
	if rc == C.PTTS_ECANCELED {
		return nil, have, true, ErrCanceled
	} else if rc != 0 {
		var buf [512]C.char
		C.ptts_ticket_error(e.h, t.id, &buf[0], C.int32_t(len(buf)))
		return nil, have, true, errors.New(C.GoString(&buf[0]))
	}
	return t.pcm[have*C.PTTS_SAMPLES_PER_FRAME : int(r)*C.PTTS_SAMPLES_PER_FRAME], int(r), d != 0, nil
}
