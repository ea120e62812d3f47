"""Device-timed stage profile of the native-cuda backend: the reference's `bench/stageprof`
(internal/bench/stageprof/stageprof.go:137-162 prints audio_ms, avg_prepare/generate/encode_ms,
avg_total_ms, rtf and the stage shares) with `generate` split into the CUDA-event stages of
ptts_stage_times_get (prefill / AR step / flow head / Mimi decode / PCM D2H).

Text preparation and the tokenizer stay in Go (out of scope here), so the input is a word count: the
chunks are the synthetic ones of SURVEY §8d (1.4 tokens per word, <= 50 tokens per chunk, forced-EOS
schedule), generated one chunk after the other like stageprof.runOnce does.

Usage (under gpurun):  python tools/stageprof.py [--words 200] [--runs 5] [--warmup 1] [--stream N]
"""
import argparse
import importlib
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge

ap = argparse.ArgumentParser()
ap.add_argument("--words", type=int, default=200)
ap.add_argument("--runs", type=int, default=5)
ap.add_argument("--warmup", type=int, default=1)
ap.add_argument("--stream", type=int, default=0, help="PTTS_REQ_STREAM with this many frames per delivery")
ap.add_argument("--arch", default="full", choices=["full", "small"])
args = ap.parse_args()

pkg = ge.package()
ck = importlib.import_module("pockettts_b200.checkpoint")
wl = importlib.import_module("pockettts_b200.workload")
import numpy as np

t0 = time.perf_counter()
tensors = ck.make_checkpoint(ck.ArchConfig.full() if args.arch == "full" else ck.ArchConfig.small())
rt = pkg.NativeCudaRuntime(tensors=tensors, max_batch=8, max_seq_tokens=512, max_frames=256, mimi_group_frames=256)
load_ms = (time.perf_counter() - t0) * 1e3
rng = np.random.default_rng([7, args.words])
chunks = wl.utterance_chunks(0, args.words, rng)


def run_once():
    t_start = time.perf_counter()
    pcm = []
    for c in chunks:  # one chunk after the other, as stageprof.runOnce
        cfg = pkg.RuntimeGenerateConfig(max_steps=c.max_steps, eos_threshold=1e30, force_eos_step=c.force_eos_step,
                                        frames_after_eos=c.frames_after_eos, noise=c.noise(), stream_frames=args.stream)
        pcm.append(rt.generate_audio(c.tokens, cfg))
    gen = time.perf_counter() - t_start
    t1 = time.perf_counter()
    audio = np.concatenate(pcm)
    np.clip(audio, -1, 1, out=audio)
    (audio * 32767).astype(np.int16).tobytes()  # the WAV payload (audio/encode.go:14-31)
    enc = time.perf_counter() - t1
    return gen, enc, audio.size


for _ in range(args.warmup):
    run_once()
rt.stage_times(reset=True)
gen = enc = 0.0
samples = 0
for _ in range(args.runs):
    g, e, samples = run_once()
    gen += g
    enc += e
st = rt.stage_times()
n = float(args.runs)
audio_ms = samples * 1000.0 / 24000.0
avg_gen, avg_enc = gen * 1e3 / n, enc * 1e3 / n
avg_total = avg_gen + avg_enc
print(f"words: {args.words}")
print(f"runs: {args.runs} (warmup {args.warmup})")
print(f"backend: native-cuda ({pkg.lib().ptts_version().decode()})")
print(f"model_load_ms: {load_ms:.2f}")
print(f"chunks: {len(chunks)}")
print(f"audio_ms: {audio_ms:.2f}")
print("avg_prepare_ms: 0.00 (text preparation stays in Go)")
print(f"avg_generate_ms: {avg_gen:.2f}")
print(f"avg_encode_ms: {avg_enc:.2f}")
print(f"avg_total_ms: {avg_total:.2f}")
print(f"rtf: {avg_total / audio_ms:.5f}")
print(f"share_generate_pct: {100 * avg_gen / avg_total:.2f}")
print(f"share_encode_pct: {100 * avg_enc / avg_total:.2f}")
print("device stages inside generate (CUDA events, per run):")
for k in ("prefill_ms", "step_ms", "flow_head_ms", "mimi_ms", "d2h_ms"):
    print(f"  {k}: {st[k] / n:.3f}")
print(f"  ar_frames: {st['steps'] / n:.0f}  prefill_tokens: {st['prefill_tokens'] / n:.0f}  "
      f"mimi_frames_decoded: {st['mimi_frames'] / n:.0f}  kernel_launches: {st['kernel_launches'] / n:.0f}")
print(f"  (flow_head_ms is part of step_ms; the B=1 frame is one fused kernel, so its head is not split out)")
rt.close()
