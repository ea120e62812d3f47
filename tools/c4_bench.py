"""C4 (BASELINE.json configs[3]): 256 voice-conditioned utterances of U{5..60} words, every request
carrying the same exported audio_prompt embedding [1,125,1024] (sin(0.01 i), as
native_safetensors_integration_test.go:81-84), continuously batched on one B200.
Prints audio-s/s and the prefill share, with and without the shared voice prefix
(PTTS_NO_VOICE_SHARING=1 in the environment disables it).  Usage (under gpurun): python tools/c4_bench.py"""
import os, sys, importlib, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge
pkg = ge.package()
ck = importlib.import_module("pockettts_b200.checkpoint")
wl = importlib.import_module("pockettts_b200.workload")
import numpy as np
import torch
rt = pkg.NativeCudaRuntime(tensors=ck.make_checkpoint(ck.ArchConfig.full()), max_batch=1024, max_seq_tokens=640,
                           max_frames=256, mimi_group_frames=2048)
TV = 125
voice = np.sin(0.01 * np.arange(TV * 1024, dtype=np.float64)).astype(np.float32).reshape(1, TV, 1024)
ve = pkg.VoiceEmbedding(voice)
chunks = wl.make_workload(256, 5, 60, seed=11)
noises = [c.noise() for c in chunks]


def run_pass():
    pend = []
    for c, n in zip(chunks, noises):
        cfg = pkg.RuntimeGenerateConfig(max_steps=c.max_steps, eos_threshold=1e30, force_eos_step=c.force_eos_step,
                                        frames_after_eos=c.frames_after_eos, noise=n, voice_embedding=ve)
        pend.append(rt.submit(c.tokens, cfg))
    return sum(rt.wait(p).n_frames for p in pend)


for _ in range(2):
    run_pass()
rt.stage_times(reset=True)
torch.cuda.synchronize()
t0 = time.perf_counter()
frames = sum(run_pass() for _ in range(2))
torch.cuda.synchronize()
dt = time.perf_counter() - t0
st = rt.stage_times()
print(f"C4: {len(chunks)} chunks, {frames} frames in {dt * 1e3:.1f} ms -> {wl.audio_seconds(frames) / dt:.0f} audio-s/s e2e; "
      f"prefill {st['prefill_ms']:.1f} ms over {st['prefill_tokens']} rows, steps {st['step_ms']:.1f} ms, "
      f"mimi {st['mimi_ms']:.1f} ms; voice cache {rt.voice_cache_stats()}; "
      f"sharing {'off' if os.environ.get('PTTS_NO_VOICE_SHARING') else 'on'}")
rt.close()
