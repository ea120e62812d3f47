"""Synthetic request workloads C1..C5 of BASELINE.json / SURVEY.md §8d.

Text preparation, tokenisation and chunking stay in Go (internal/text,
internal/tokenizer); a benchmark therefore only needs their OUTPUT statistics:
per chunk the token count, the word count (FramesAfterEOS rule) and the frame
budget.  The rules mirrored here:
  * 1.4 tokens per word (rounded), sentences of 8..15 words, greedy packing of
    whole sentences into <= 50-token chunks (internal/text/prepare.go:105-185,
    internal/tts/service.go:23);
  * < 5 words: PrepareText left-pads 8 spaces = 8 extra tokens of id 260
    (prepare.go:92-97, tokenizer_test.go:110-137);
  * max frames = ceil((n_tok/3 + 2) * 12.5) (prepare.go:38-48);
  * FramesAfterEOS = 5 if <= 4 words else 3 (prepare.go:53-59);
  * random-init weights have no meaningful EOS logit, so EOS is declared at
    step round(n_tok * 12.5 / 3) (the reference's own 3 tokens/s assumption),
    identically for the CPU and GPU arms.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np

FRAME_RATE = 12.5
MAX_CHUNK_TOKENS = 50
PAD_TOKEN = 260


@dataclass
class Chunk:
    utt: int
    index: int
    tokens: np.ndarray
    n_words: int
    max_steps: int
    force_eos_step: int
    frames_after_eos: int
    noise_seed: int

    @property
    def expected_frames(self) -> int:
        return min(self.force_eos_step + 1 + self.frames_after_eos, self.max_steps)

    def noise(self, temperature: float = 0.7) -> np.ndarray:
        rng = np.random.default_rng(self.noise_seed)
        return (rng.standard_normal((self.max_steps, 32)) * math.sqrt(temperature)).astype(np.float32)


def _chunk(utt, index, n_tok, n_words, rng, seed):
    toks = rng.integers(0, 4000, size=n_tok).astype(np.int64)
    if n_words < 5:
        toks = np.concatenate([np.full(8, PAD_TOKEN, np.int64), toks])
    n = int(toks.size)
    max_steps = int(math.ceil((n / 3.0 + 2.0) * FRAME_RATE))
    eos = int(round(n * FRAME_RATE / 3.0))
    return Chunk(utt, index, toks, n_words, max_steps, min(eos, max_steps - 1), 5 if n_words <= 4 else 3, seed)


def utterance_chunks(utt: int, n_words: int, rng: np.random.Generator, seed_base: int = 1234):
    """Greedy whole-sentence packing into <= 50-token chunks."""
    sentences = []
    left = n_words
    while left > 0:
        w = int(min(left, rng.integers(8, 16)))
        sentences.append(w)
        left -= w
    chunks, cur_w, cur_t = [], 0, 0
    for w in sentences:
        t = max(1, int(round(1.4 * w)))
        if cur_t and cur_t + t > MAX_CHUNK_TOKENS:
            chunks.append((cur_w, cur_t))
            cur_w, cur_t = 0, 0
        cur_w += w
        cur_t += t
    if cur_t:
        chunks.append((cur_w, cur_t))
    out = []
    for i, (w, t) in enumerate(chunks):
        out.append(_chunk(utt, i, min(t, MAX_CHUNK_TOKENS), w, rng, seed_base + utt * 16 + i))
    return out


def make_workload(n_utterances: int, words_lo: int, words_hi: int, seed: int = 7, first_utt: int = 0):
    """Chunks of `n_utterances` utterances with word counts U{lo..hi}."""
    chunks = []
    for u in range(first_utt, first_utt + n_utterances):
        rng = np.random.default_rng([seed, u])
        n_words = int(rng.integers(words_lo, words_hi + 1))
        chunks.extend(utterance_chunks(u, n_words, rng))
    return chunks


def hello_world() -> Chunk:
    """C1: "Hello world." -> 8 pad tokens + [2994, 578, 263] (tokenizer_test.go:110-137)."""
    toks = np.array([PAD_TOKEN] * 8 + [2994, 578, 263], np.int64)
    n = toks.size
    max_steps = int(math.ceil((n / 3.0 + 2.0) * FRAME_RATE))
    return Chunk(0, 0, toks, 2, max_steps, int(round(n * FRAME_RATE / 3.0)), 5, 1234)


def paragraph_200_words(seed: int = 7):
    """C2: one 200-word paragraph -> ~6 chunks, streamed in order."""
    rng = np.random.default_rng([seed, 200])
    return utterance_chunks(0, 200, rng)


def audio_seconds(frames: int) -> float:
    return frames * 1920 / 24000.0
