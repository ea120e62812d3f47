//go:build cgo && linux

package tts

import (
	"context"
	"errors"
	"fmt"
	"log/slog"
	"math"
	"math/rand"
	"sync"
	"time"

	"github.com/cwbudde/go-pocket-tts/internal/native/cuda"
	"github.com/cwbudde/go-pocket-tts/internal/text"
)

// nativeCUDARuntime is the B200 backend behind tts.Runtime.  It keeps the
// contract of nativeSafetensorsRuntime (runtime_native_safetensors.go:52-238):
// same argument meaning, same guard strings, same EOS-countdown rule (applied
// inside the engine, per request), fresh PCM slice per call, goroutine-safe.
// What changes: concurrent callers are batched by the engine instead of being
// serialised by rngMu, and the noise is drawn here and passed down.
type nativeCUDARuntime struct {
	engine *cuda.Engine

	rngMu sync.Mutex
	rng   *rand.Rand
}

func newNativeCUDARuntime(engine *cuda.Engine) Runtime {
	return &nativeCUDARuntime{engine: engine, rng: rand.New(rand.NewSource(time.Now().UnixNano()))}
}

func (r *nativeCUDARuntime) MimiTiming() (float64, float64, int) { return 12.5, 200, 16 }

func (r *nativeCUDARuntime) GenerateAudio(ctx context.Context, tokens []int64, cfg RuntimeGenerateConfig) ([]float32, error) {
	if r == nil || r.engine == nil {
		return nil, errors.New("native-cuda runtime unavailable")
	}
	if len(tokens) == 0 {
		return nil, errors.New("generate: token slice must not be empty")
	}
	if cfg.VoiceEmbedding != nil && cfg.VoiceModelState != nil {
		return nil, errors.New("generate: voice embedding and voice model state are mutually exclusive")
	}

	maxSteps := cfg.MaxSteps
	if maxSteps <= 0 {
		maxSteps = cfg.EstimatedMaxSteps
	}
	if maxSteps <= 0 {
		maxSteps = text.EstimateMaxFrames(len(tokens), text.DefaultMimiFrameRate)
	}

	// makeGaussianNoise (flow_lm.go:386-408), hoisted out of the step: one row
	// of 32 per potential frame.
	sigma := math.Sqrt(math.Max(cfg.Temperature, 0))
	noise := make([]float32, maxSteps*32)
	r.rngMu.Lock()
	for i := range noise {
		noise[i] = float32(r.rng.NormFloat64() * sigma)
	}
	r.rngMu.Unlock()

	var voiceEmb []float32
	voiceFrames := 0
	if cfg.VoiceEmbedding != nil {
		voiceEmb = cfg.VoiceEmbedding.Data
		if n := len(cfg.VoiceEmbedding.Shape); n >= 2 {
			voiceFrames = int(cfg.VoiceEmbedding.Shape[n-2])
		}
	}

	var voice []cuda.VoiceLayer
	if cfg.VoiceModelState != nil {
		for i := 0; ; i++ {
			module := cfg.VoiceModelState.Modules[fmt.Sprintf("transformer.layers.%d.self_attn", i)]
			if module == nil {
				break
			}
			cache, offset := module["cache"], module["offset"]
			if cache == nil || offset == nil || len(cache.Shape) != 5 || len(offset.Data) == 0 {
				return nil, fmt.Errorf("generate: load voice model state: module %d malformed", i)
			}
			voice = append(voice, cuda.VoiceLayer{Cache: cache.Data, CacheTokens: int(cache.Shape[2]), Offset: int(offset.Data[0])})
		}
	}

	start := time.Now()
	ticket, err := r.engine.Submit(tokens, voiceEmb, voiceFrames, voice, cuda.GenConfig{
		Temperature: cfg.Temperature, EOSThreshold: cfg.EOSThreshold, MaxSteps: maxSteps,
		EstimatedMaxSteps: cfg.EstimatedMaxSteps, LSDDecodeSteps: cfg.LSDDecodeSteps, FramesAfterEOS: cfg.FramesAfterEOS,
	}, noise, maxSteps)
	if err != nil {
		return nil, err
	}

	// ctx cancellation (reference checks ctx.Err() before every step, :156-159).
	done := make(chan struct{})
	go func() {
		select {
		case <-ctx.Done():
			r.engine.Cancel(ticket)
		case <-done:
		}
	}()
	pcm, frames, err := r.engine.Wait(ticket)
	close(done)
	if errors.Is(err, cuda.ErrCanceled) {
		return nil, ctx.Err()
	}
	if err != nil {
		return nil, err
	}

	if cfg.StepCallback != nil {
		for s := 1; s <= frames; s++ {
			cfg.StepCallback(s, maxSteps)
		}
	}
	slog.Info("generation complete", "backend", "native-cuda", "frames", frames, "samples", len(pcm),
		"duration_ms", time.Since(start).Milliseconds())
	return pcm, nil
}

func (r *nativeCUDARuntime) Close() {
	if r != nil && r.engine != nil {
		r.engine.Close()
	}
}
