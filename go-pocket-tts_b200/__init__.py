"""pockettts-b200: B200-native `native-cuda` synthesis backend for go-pocket-tts.

This package is the host-side mirror of the reference's `tts.Runtime` interface
(/root/reference internal/tts/runtime.go:42-45) on top of the C ABI declared in
include/pockettts_cuda.h.  It exists so the parity tests and the benchmark can
drive the CUDA library the way the Go backend would; the product is the shared
library `libpockettts_cuda.so` (sm_100a kernels + engine + scheduler).

There is no CPU fallback: if the CUDA library is missing or no sm_100 device is
present, construction fails loudly.  (The directory name contains hyphens, so it
is loaded through `__graft_entry__.package()` as module `pockettts_b200`.)
"""
from __future__ import annotations

import ctypes as C
import math
import os
import subprocess
from dataclasses import dataclass, field
from typing import Callable, Optional, Sequence

import numpy as np

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
REPO_ROOT = os.path.dirname(PKG_DIR)
LIB_PATH = os.path.join(PKG_DIR, "libpockettts_cuda.so")
CSRC = ["gemm.cu", "tcgemm.cu", "attention.cu", "attention_decode.cu", "attention_tc.cu", "attention_mma.cu", "step_fused.cu", "step_ring.cu", "flow_head.cu", "resblock.cu", "elementwise.cu", "engine.cu", "scheduler.cu", "pool.cu", "safetensors.cpp"]

PTTS_OK, PTTS_EINVAL, PTTS_ECUDA, PTTS_ENOMEM, PTTS_ESTATE, PTTS_ECANCELED = 0, -1, -2, -3, -4, -5
PTTS_REQ_NO_PCM_D2H = 1
PTTS_REQ_PCM16 = 2
PTTS_REQ_STREAM = 4
LATENT_DIM = 32
SAMPLES_PER_FRAME = 1920
SAMPLE_RATE = 24000


class NativeCudaError(RuntimeError):
    def __init__(self, status: int, msg: str):
        super().__init__(msg)
        self.status = status


# ---------------------------------------------------------------------------
# build
# ---------------------------------------------------------------------------
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC,-ffp-contract=off"]
OBJ_DIR = os.path.join(PKG_DIR, "csrc", "_obj")


def nvcc_command(out: str = LIB_PATH) -> list[str]:
    """The one-shot equivalent of what build_library does per file."""
    srcs = [os.path.join(PKG_DIR, "csrc", s) for s in CSRC]
    return ["nvcc", *NVCC_FLAGS, "-shared", "-o", out, *srcs]


def build_library(force: bool = False, verbose: bool = False) -> str:
    """Compile every CUDA source for sm_100a into the in-tree shared library: one object per
    source (rebuilt when the source or any header is newer), compiled in parallel, then linked."""
    from concurrent.futures import ThreadPoolExecutor
    cdir = os.path.join(PKG_DIR, "csrc")
    headers = [os.path.join(cdir, h) for h in os.listdir(cdir) if h.endswith((".cuh", ".h", ".hpp"))]
    headers.append(os.path.join(REPO_ROOT, "include", "pockettts_cuda.h"))
    hdr_time = max(os.path.getmtime(h) for h in headers)
    src_time = max(os.path.getmtime(os.path.join(cdir, n)) for n in CSRC)
    if not force and os.path.exists(LIB_PATH) and os.path.getmtime(LIB_PATH) >= max(hdr_time, src_time):
        return LIB_PATH  # e.g. on the GPU box: the built library travels, the objects do not
    os.makedirs(OBJ_DIR, exist_ok=True)
    jobs, objs = [], []
    for name in CSRC:
        src = os.path.join(cdir, name)
        obj = os.path.join(OBJ_DIR, name + ".o")
        objs.append(obj)
        if force or not os.path.exists(obj) or os.path.getmtime(obj) < max(os.path.getmtime(src), hdr_time):
            jobs.append(["nvcc", *NVCC_FLAGS, "-c", "-o", obj, src])
    def run(cmd):
        if verbose:
            print(" ".join(cmd), flush=True)
        subprocess.run(cmd, check=True)

    with ThreadPoolExecutor(max_workers=max(1, min(len(jobs), os.cpu_count() or 1))) as ex:
        list(ex.map(run, jobs))
    run(["nvcc", *NVCC_FLAGS, "-shared", "-o", LIB_PATH, *objs])
    return LIB_PATH


# ---------------------------------------------------------------------------
# ctypes mirror of include/pockettts_cuda.h
# ---------------------------------------------------------------------------
class PttsConfig(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("device", "max_batch", "page_tokens", "max_seq_tokens", "kv_pages",
                                         "max_frames", "mimi_group_frames", "max_prefill_tokens", "flags")]


class PttsGenCfg(C.Structure):
    _fields_ = [("temperature", C.c_double), ("eos_threshold", C.c_double), ("max_steps", C.c_int32),
                ("estimated_max_steps", C.c_int32), ("lsd_decode_steps", C.c_int32),
                ("frames_after_eos", C.c_int32), ("force_eos_step", C.c_int32), ("flags", C.c_int32),
                ("stream_frames", C.c_int32)]


class PttsVoiceLayer(C.Structure):
    _fields_ = [("cache", C.POINTER(C.c_float)), ("cache_tokens", C.c_int32), ("offset", C.c_int32)]


class PttsVoiceKV(C.Structure):
    _fields_ = [("layers", C.POINTER(PttsVoiceLayer)), ("n_layers", C.c_int32)]


class PttsModelInfo(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "d_model", "n_heads", "head_dim", "n_layers", "ffn_dim", "flow_dim", "flow_depth", "latent_dim",
        "n_bins", "mimi_dim", "mimi_heads", "mimi_layers", "mimi_ffn_dim", "mimi_context",
        "samples_per_frame", "mimi_steps_per_latent")] + [(n, C.c_int64) for n in (
            "params_ar", "params_mimi", "macs_ar_frame", "macs_mimi_frame", "macs_prefill_token")]


class PttsStageTimes(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("prefill_ms", "step_ms", "flow_head_ms", "mimi_ms", "d2h_ms")] + \
               [(n, C.c_int64) for n in ("prefill_tokens", "step_rows", "steps", "mimi_frames", "kernel_launches")]


class PttsResult(C.Structure):
    _fields_ = [("status", C.c_int32), ("n_frames", C.c_int32), ("n_samples", C.c_int64),
                ("queue_ms", C.c_double), ("first_frame_ms", C.c_double), ("total_ms", C.c_double)]


EXPORTS = [
    "ptts_last_error", "ptts_version", "ptts_device_count", "ptts_create", "ptts_destroy", "ptts_load_tensor",
    "ptts_load_safetensors", "ptts_finalize_weights", "ptts_model_info_get", "ptts_generate_audio",
    "ptts_submit", "ptts_wait", "ptts_cancel", "ptts_ticket_latents", "ptts_ticket_error", "ptts_ticket_read",
    "ptts_ticket_release", "ptts_seq_open",
    "ptts_step", "ptts_decode", "ptts_decode_latents", "ptts_mimi_encode_to_latent", "ptts_seq_frames", "ptts_seq_kv_len", "ptts_seq_read_kv",
    "ptts_seq_close", "ptts_stage_times_get", "ptts_debug_fetch", "ptts_profile_enable", "ptts_profile_report",
    "ptts_voice_cache_stats", "ptts_debug_gemm", "ptts_debug_gemm_bench", "ptts_debug_tc_force", "ptts_debug_gemm_tma",
    "ptts_pool_create", "ptts_pool_destroy", "ptts_pool_size", "ptts_pool_engine", "ptts_pool_load_tensor",
    "ptts_pool_load_safetensors", "ptts_pool_finalize_weights", "ptts_pool_submit", "ptts_pool_wait", "ptts_pool_cancel",
    "ptts_pool_ticket_error", "ptts_pool_ticket_latents", "ptts_pool_ticket_read", "ptts_pool_ticket_release", "ptts_pool_generate_audio",
    "ptts_pool_load",
]

_lib = None


def lib() -> C.CDLL:
    """Loads the CUDA library; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise NativeCudaError(PTTS_ESTATE, f"native-cuda runtime unavailable: {LIB_PATH} not built "
                                           f"(run __graft_entry__.build())")
    L = C.CDLL(LIB_PATH)
    fp, ip, i64p = C.POINTER(C.c_float), C.POINTER(C.c_int32), C.POINTER(C.c_int64)
    vp = C.c_void_p
    L.ptts_last_error.restype = C.c_char_p
    L.ptts_version.restype = C.c_char_p
    L.ptts_device_count.restype = C.c_int
    L.ptts_create.argtypes = [C.POINTER(PttsConfig), C.POINTER(vp)]
    L.ptts_destroy.argtypes = [vp]
    L.ptts_destroy.restype = None
    L.ptts_load_tensor.argtypes = [vp, C.c_char_p, i64p, C.c_int32, C.c_int32, vp]
    L.ptts_load_safetensors.argtypes = [vp, C.c_char_p]
    L.ptts_finalize_weights.argtypes = [vp]
    L.ptts_model_info_get.argtypes = [vp, C.POINTER(PttsModelInfo)]
    gen_args = [vp, i64p, C.c_int32, fp, C.c_int32, C.POINTER(PttsVoiceKV), C.POINTER(PttsGenCfg), fp,
                C.c_int32, fp, C.c_int64]
    L.ptts_generate_audio.argtypes = gen_args + [C.POINTER(PttsResult)]
    L.ptts_submit.argtypes = gen_args + [i64p]
    L.ptts_wait.argtypes = [vp, C.c_int64, C.POINTER(PttsResult)]
    L.ptts_cancel.argtypes = [vp, C.c_int64]
    L.ptts_ticket_latents.argtypes = [vp, C.c_int64, fp, fp, C.c_int32]
    L.ptts_ticket_error.argtypes = [vp, C.c_int64, C.c_char_p, C.c_int32]
    L.ptts_ticket_read.argtypes = [vp, C.c_int64, C.c_int32, C.c_int32, ip, ip]
    L.ptts_ticket_release.argtypes = [vp, C.c_int64]
    L.ptts_seq_open.argtypes = [vp, i64p, C.c_int32, fp, C.c_int32, C.POINTER(PttsVoiceKV), C.c_int32,
                                C.POINTER(vp)]
    L.ptts_step.argtypes = [vp, C.POINTER(vp), C.c_int32, fp, fp, C.c_int32, fp, fp]
    L.ptts_decode.argtypes = [vp, vp, C.c_int32, C.c_int32, fp]
    L.ptts_decode_latents.argtypes = [vp, fp, ip, C.c_int32, fp]
    L.ptts_seq_frames.argtypes = [vp]
    L.ptts_seq_kv_len.argtypes = [vp]
    L.ptts_seq_read_kv.argtypes = [vp, vp, C.c_int32, C.c_int32, fp]
    L.ptts_seq_close.argtypes = [vp, vp]
    L.ptts_stage_times_get.argtypes = [vp, C.POINTER(PttsStageTimes), C.c_int32]
    L.ptts_debug_fetch.argtypes = [vp, C.c_char_p, fp, C.c_int64]
    L.ptts_debug_fetch.restype = C.c_int64
    L.ptts_profile_enable.argtypes = [vp, C.c_int32]
    L.ptts_debug_tc_force.argtypes = [C.c_int32, C.c_int32]
    L.ptts_voice_cache_stats.argtypes = [vp, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_int32)]
    L.ptts_debug_gemm_tma.argtypes = [vp, fp, C.c_int32, C.c_int32, C.c_int32, fp, C.c_int32, fp, C.c_int32, fp, fp]
    L.ptts_debug_gemm_bench.argtypes = [vp, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, fp]
    L.ptts_debug_gemm.argtypes = [vp, fp, C.c_int32, C.c_int32, fp, C.c_int32, fp, C.c_int32, C.c_int32, fp]
    L.ptts_profile_report.argtypes = [vp, C.c_char_p, C.c_int32]
    L.ptts_pool_create.argtypes = [C.POINTER(PttsConfig), ip, C.c_int32, C.POINTER(vp)]
    L.ptts_pool_destroy.argtypes = [vp]
    L.ptts_pool_destroy.restype = None
    L.ptts_pool_size.argtypes = [vp]
    L.ptts_pool_engine.argtypes = [vp, C.c_int32, C.POINTER(vp)]
    L.ptts_pool_load_tensor.argtypes = [vp, C.c_char_p, i64p, C.c_int32, C.c_int32, vp]
    L.ptts_pool_load_safetensors.argtypes = [vp, C.c_char_p]
    L.ptts_pool_finalize_weights.argtypes = [vp]
    L.ptts_pool_submit.argtypes = gen_args + [i64p, ip]
    L.ptts_pool_wait.argtypes = [vp, C.c_int64, C.POINTER(PttsResult)]
    L.ptts_pool_cancel.argtypes = [vp, C.c_int64]
    L.ptts_pool_ticket_error.argtypes = [vp, C.c_int64, C.c_char_p, C.c_int32]
    L.ptts_pool_ticket_latents.argtypes = [vp, C.c_int64, fp, fp, C.c_int32]
    L.ptts_pool_ticket_read.argtypes = [vp, C.c_int64, C.c_int32, C.c_int32, ip, ip]
    L.ptts_pool_ticket_release.argtypes = [vp, C.c_int64]
    L.ptts_pool_generate_audio.argtypes = gen_args + [C.POINTER(PttsResult)]
    L.ptts_pool_load.argtypes = [vp, i64p, ip, i64p, C.c_int32]
    _lib = L
    return L


def _fptr(a: Optional[np.ndarray]):
    if a is None:
        return None
    assert a.dtype == np.float32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _check(rc: int):
    if rc != 0:
        raise NativeCudaError(rc, lib().ptts_last_error().decode("utf-8", "replace"))


# ---------------------------------------------------------------------------
# reference-shaped configuration objects
# ---------------------------------------------------------------------------
@dataclass
class VoiceEmbedding:
    """tts.VoiceEmbedding (internal/tts/runtime.go:11-15): data [1,T,D] or [T,D]."""
    data: np.ndarray

    def rows(self) -> np.ndarray:
        a = np.ascontiguousarray(self.data, dtype=np.float32)
        if a.ndim == 3:
            if a.shape[0] != 1:
                raise ValueError("voice embedding batch must be 1")
            a = a[0]
        if a.ndim != 2:
            raise ValueError(f"voice embedding has {a.ndim}D shape, expected 2D or 3D")
        return np.ascontiguousarray(a)


@dataclass
class VoiceModelState:
    """safetensors.VoiceModelState: per-layer `cache` [2,1,T,H,Dh] + `offset`
    (internal/safetensors/reader.go:273-308, flow_transformer.go:451-480)."""
    caches: list  # list of np.ndarray [2,1,T,H,Dh]
    offsets: list  # list of int


@dataclass
class RuntimeGenerateConfig:
    """tts.RuntimeGenerateConfig (internal/tts/runtime.go:17-31).  `noise` is the
    host-drawn N(0, temperature) array [steps, 32] (flow_lm.go:386-408 draws it
    from math/rand inside the step; here it is an explicit input)."""
    temperature: float = 0.7
    eos_threshold: float = -4.0
    max_steps: int = 0
    estimated_max_steps: int = 0
    lsd_decode_steps: int = 1
    frames_after_eos: int = 3
    voice_embedding: Optional[VoiceEmbedding] = None
    voice_model_state: Optional[VoiceModelState] = None
    step_callback: Optional[Callable[[int, int], None]] = None
    noise: Optional[np.ndarray] = None
    seed: int = 0
    force_eos_step: int = -1
    no_pcm_d2h: bool = False
    pcm16: bool = False  # PCM16 samples quantised on the device (wav_stream.go:44-54); GenerateResult.pcm is int16
    stream_frames: int = 0  # > 0: frame-level streaming (PTTS_REQ_STREAM), PCM delivered every this many frames
    # Caller-owned output buffer (>= max steps * 1920 samples, float32 or int16 with pcm16), e.g. one a server
    # reuses across requests: the C ABI writes PCM into whatever `pcm_out` it is given (ptts_submit).
    pcm_out: Optional[np.ndarray] = None


def estimate_max_frames(n_tokens: int, frame_rate: float = 12.5) -> int:
    """internal/text/prepare.go:38-48."""
    n_tokens = max(n_tokens, 0)
    if not (frame_rate > 0) or math.isinf(frame_rate):
        frame_rate = 12.5
    return int(math.ceil((n_tokens / 3.0 + 2.0) * frame_rate))


def resolve_max_steps(cfg: RuntimeGenerateConfig, n_tokens: int) -> int:
    """internal/tts/runtime_native_safetensors.go:61-67."""
    ms = cfg.max_steps
    if ms <= 0:
        ms = cfg.estimated_max_steps
    if ms <= 0:
        ms = estimate_max_frames(n_tokens)
    return ms


def make_noise(steps: int, temperature: float, seed: int) -> np.ndarray:
    """Host noise: N(0,1)*sqrt(max(temperature,0)) as float32 (flow_lm.go:386-408)."""
    sigma = math.sqrt(max(float(temperature), 0.0))
    rng = np.random.default_rng(seed)
    return (rng.standard_normal((steps, LATENT_DIM)) * sigma).astype(np.float32)


@dataclass
class GenerateResult:
    pcm: np.ndarray
    n_frames: int
    queue_ms: float = 0.0
    first_frame_ms: float = 0.0
    total_ms: float = 0.0
    latents: Optional[np.ndarray] = None
    eos_logits: Optional[np.ndarray] = None


class _Pending:
    def __init__(self, ticket, keep, pcm, max_steps):
        self.ticket, self.keep, self.pcm, self.max_steps = ticket, keep, pcm, max_steps


class NativeCudaRuntime:
    """`tts.Runtime` for --backend native-cuda.

    Mirrors nativeSafetensorsRuntime (internal/tts/runtime_native_safetensors.go):
    `generate_audio(tokens, cfg) -> float32 PCM`, `close()` (nil-safe, idempotent),
    `mimi_timing()`.  `generate_audio` may be called from many threads; the engine
    batches them continuously."""

    def __init__(self, checkpoint: Optional[str] = None, *, device: int = 0, max_batch: int = 256,
                 max_seq_tokens: int = 1024, max_frames: int = 512, mimi_group_frames: int = 0,
                 kv_pages: int = 0, tensors: Optional[dict] = None, flags: int = 0):
        self._lib = lib()
        self._h = C.c_void_p()
        cfg = PttsConfig(device=device, max_batch=max_batch, max_seq_tokens=max_seq_tokens,
                         max_frames=max_frames, mimi_group_frames=mimi_group_frames, kv_pages=kv_pages,
                         flags=flags)
        _check(self._lib.ptts_create(C.byref(cfg), C.byref(self._h)))
        try:
            if checkpoint is not None:
                _check(self._lib.ptts_load_safetensors(self._h, checkpoint.encode()))
            if tensors:
                for k, a in tensors.items():
                    self.load_tensor(k, a)
            if checkpoint is not None or tensors:
                self.finalize()
        except Exception:
            self.close()
            raise

    # -- weights ----------------------------------------------------------
    def load_tensor(self, key: str, a: np.ndarray, dtype: str = "f32"):
        """One safetensors key.  dtype "f16" / "bf16" hands the engine the half-precision payload
        (PTTS_DTYPE_F16 / _BF16), widened on the C side as store.go:339-395 does."""
        a = np.ascontiguousarray(a, dtype=np.float32)
        shape = (C.c_int64 * a.ndim)(*a.shape)
        code = {"f32": 0, "f16": 1, "bf16": 2}[dtype]
        if code == 1:
            a = np.ascontiguousarray(a.astype(np.float16))
        elif code == 2:
            a = np.ascontiguousarray((a.view(np.uint32) >> 16).astype(np.uint16))  # exact for bf16-representable values
        _check(self._lib.ptts_load_tensor(self._h, key.encode(), shape, a.ndim, code, a.ctypes.data_as(C.c_void_p)))

    def finalize(self):
        _check(self._lib.ptts_finalize_weights(self._h))
        info = PttsModelInfo()
        _check(self._lib.ptts_model_info_get(self._h, C.byref(info)))
        self.info = info

    # -- tts.Runtime ------------------------------------------------------
    def close(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            self._lib.ptts_destroy(h)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def mimi_timing(self):
        """(frame_rate, encoder_frame_rate, steps_per_latent), service.go:280-298."""
        return 12.5, 200.0, 16

    def _voice_args(self, cfg: RuntimeGenerateConfig, keep: list):
        if cfg.voice_embedding is not None and cfg.voice_model_state is not None:
            raise NativeCudaError(PTTS_EINVAL,
                                  "generate: voice embedding and voice model state are mutually exclusive")
        vemb, vframes, vkv = None, 0, None
        if cfg.voice_embedding is not None:
            rows = cfg.voice_embedding.rows()
            keep.append(rows)
            vemb, vframes = _fptr(rows), rows.shape[0]
        if cfg.voice_model_state is not None:
            st = cfg.voice_model_state
            n = len(st.caches)
            arr = (PttsVoiceLayer * n)()
            for i, (c, off) in enumerate(zip(st.caches, st.offsets)):
                c = np.ascontiguousarray(c, dtype=np.float32)
                if c.ndim != 5 or c.shape[0] != 2 or c.shape[1] != 1:
                    raise NativeCudaError(PTTS_EINVAL,
                                          f"native: voice model state cache shape {list(c.shape)}, want [2,B,T,H,D]")
                keep.append(c)
                arr[i].cache = _fptr(c)
                arr[i].cache_tokens = c.shape[2]
                arr[i].offset = int(off)
            kv = PttsVoiceKV(layers=arr, n_layers=n)
            keep.extend([arr, kv])
            vkv = C.byref(kv)
        return vemb, vframes, vkv

    def submit(self, tokens: Sequence[int], cfg: RuntimeGenerateConfig) -> _Pending:
        if self._h is None:
            raise NativeCudaError(PTTS_ESTATE, "native-cuda runtime unavailable")
        toks = np.ascontiguousarray(np.asarray(tokens, dtype=np.int64))
        if toks.size == 0:
            raise NativeCudaError(PTTS_EINVAL, "generate: token slice must not be empty")
        ms = resolve_max_steps(cfg, toks.size)
        noise = cfg.noise
        if noise is None:
            noise = make_noise(ms, cfg.temperature, cfg.seed)
        noise = np.ascontiguousarray(noise, dtype=np.float32)
        keep: list = [toks, noise]
        vemb, vframes, vkv = self._voice_args(cfg, keep)
        flags = PTTS_REQ_NO_PCM_D2H if cfg.no_pcm_d2h else 0
        if getattr(cfg, "pcm16", False):
            flags |= PTTS_REQ_PCM16
        if getattr(cfg, "stream_frames", 0) > 0:
            flags |= PTTS_REQ_STREAM
        g = PttsGenCfg(temperature=cfg.temperature, eos_threshold=cfg.eos_threshold, max_steps=cfg.max_steps,
                       estimated_max_steps=cfg.estimated_max_steps, lsd_decode_steps=cfg.lsd_decode_steps,
                       frames_after_eos=cfg.frames_after_eos, force_eos_step=cfg.force_eos_step, flags=flags,
                       stream_frames=max(getattr(cfg, "stream_frames", 0), 0))
        want = np.int16 if getattr(cfg, "pcm16", False) else np.float32
        pcm = None
        if not cfg.no_pcm_d2h:
            pcm = getattr(cfg, "pcm_out", None)
            if pcm is None:
                pcm = np.empty(ms * SAMPLES_PER_FRAME, dtype=want)
            elif (pcm.dtype != want or pcm.ndim != 1 or pcm.size < ms * SAMPLES_PER_FRAME or
                  not pcm.flags["C_CONTIGUOUS"] or not pcm.flags["WRITEABLE"]):
                raise NativeCudaError(PTTS_EINVAL, "generate: pcm_out must be a writable contiguous 1-D %s array of "
                                                   ">= %d samples" % (np.dtype(want).name, ms * SAMPLES_PER_FRAME))
        ticket = C.c_int64()
        _check(self._submit(toks.ctypes.data_as(C.POINTER(C.c_int64)), toks.size, vemb, vframes,
                            vkv, C.byref(g), _fptr(noise), noise.shape[0],
                            None if pcm is None else pcm.ctypes.data_as(C.POINTER(C.c_float)),
                            0 if pcm is None else pcm.size, C.byref(ticket)))
        return _Pending(ticket.value, keep, pcm, ms)

    # the five entry points a pool overrides (ptts_pool_*)
    def _submit(self, *a):
        return self._lib.ptts_submit(self._h, *a)

    def _wait(self, ticket, res):
        return self._lib.ptts_wait(self._h, ticket, res)

    def _ticket_error(self, ticket, buf, cap):
        return self._lib.ptts_ticket_error(self._h, ticket, buf, cap)

    def _ticket_latents(self, ticket, lat, eos, cap):
        return self._lib.ptts_ticket_latents(self._h, ticket, lat, eos, cap)

    def _ticket_release(self, ticket):
        return self._lib.ptts_ticket_release(self._h, ticket)

    def _ticket_read(self, ticket, have, timeout, ready, done):
        return self._lib.ptts_ticket_read(self._h, ticket, have, timeout, ready, done)

    def _cancel(self, ticket):
        return self._lib.ptts_cancel(self._h, ticket)

    def wait(self, p: _Pending, want_latents: bool = False) -> GenerateResult:
        res = PttsResult()
        rc = self._wait(p.ticket, C.byref(res))
        try:
            if rc != 0:  # the request's own message (not the calling thread's last error)
                if self._h is None:  # engine closed under the waiter: ptts_wait left the text on this thread
                    raise NativeCudaError(rc, self._lib.ptts_last_error().decode("utf-8", "replace"))
                buf = C.create_string_buffer(512)
                self._ticket_error(p.ticket, buf, len(buf))
                raise NativeCudaError(rc, buf.value.decode("utf-8", "replace"))
            lat = eos = None
            if want_latents:
                lat = np.empty((res.n_frames, LATENT_DIM), dtype=np.float32)
                eos = np.empty(res.n_frames, dtype=np.float32)
                n = self._ticket_latents(p.ticket, _fptr(lat), _fptr(eos), res.n_frames)
                if n < 0:
                    _check(n)
            pcm = p.pcm[: res.n_samples] if p.pcm is not None else np.empty(0, np.float32)  # view: no extra host copy
            return GenerateResult(pcm, res.n_frames, res.queue_ms, res.first_frame_ms, res.total_ms, lat, eos)
        finally:
            if self._h is not None:
                self._ticket_release(p.ticket)

    def cancel(self, p: _Pending):
        _check(self._cancel(p.ticket))

    def read(self, p: _Pending, have_frames: int = 0, timeout_ms: int = -1):
        """ptts_ticket_read: (frames_ready, done).  PCM of frames [0, frames_ready) is valid in p.pcm."""
        ready, done = C.c_int32(0), C.c_int32(0)
        rc = self._ticket_read(p.ticket, have_frames, timeout_ms, C.byref(ready), C.byref(done))
        if rc != 0:
            raise NativeCudaError(rc, self._lib.ptts_last_error().decode("utf-8", "replace"))
        return ready.value, bool(done.value)

    def stream(self, tokens, cfg: RuntimeGenerateConfig):
        """SynthesizeStream at frame granularity (tts/runtime.go:33-38 PCMChunk): yields the PCM of the
        frames delivered since the previous yield, as views into one buffer."""
        if cfg.stream_frames <= 0:
            cfg = RuntimeGenerateConfig(**{**cfg.__dict__, "stream_frames": 1})
        p = self.submit(tokens, cfg)
        have = 0
        try:
            while True:
                ready, done = self.read(p, have)
                if ready > have:
                    yield p.pcm[have * SAMPLES_PER_FRAME: ready * SAMPLES_PER_FRAME]
                    have = ready
                if done:
                    break
        finally:
            self.wait(p)

    def generate_audio(self, tokens: Sequence[int], cfg: Optional[RuntimeGenerateConfig] = None) -> np.ndarray:
        """GenerateAudio(ctx, tokens, cfg) ([]float32, error)."""
        return self.generate(tokens, cfg).pcm

    def generate(self, tokens, cfg: Optional[RuntimeGenerateConfig] = None, want_latents=False) -> GenerateResult:
        cfg = cfg or RuntimeGenerateConfig()
        r = self.wait(self.submit(tokens, cfg), want_latents)
        if cfg.step_callback is not None:  # reported after the fact: the loop runs on the device side
            ms = resolve_max_steps(cfg, len(tokens))
            for s in range(r.n_frames):
                cfg.step_callback(s + 1, ms)
        return r

    # -- stage-level API (native.Model mirror) ------------------------------
    def seq_open(self, tokens, voice_embedding: Optional[VoiceEmbedding] = None,
                 voice_model_state: Optional[VoiceModelState] = None, reserve_frames: int = 0) -> "Sequence_":
        toks = np.ascontiguousarray(np.asarray(tokens, dtype=np.int64))
        keep: list = [toks]
        cfg = RuntimeGenerateConfig(voice_embedding=voice_embedding, voice_model_state=voice_model_state)
        vemb, vframes, vkv = self._voice_args(cfg, keep)
        h = C.c_void_p()
        _check(self._lib.ptts_seq_open(self._h, toks.ctypes.data_as(C.POINTER(C.c_int64)), toks.size, vemb,
                                       vframes, vkv, reserve_frames, C.byref(h)))
        return Sequence_(self, h)

    def step(self, seqs, noise: np.ndarray, frames_in: Optional[np.ndarray] = None, lsd_steps: int = 1):
        """SampleNextLatentStateful for a batch: returns (latents [B,32], eos [B])."""
        B = len(seqs)
        noise = np.ascontiguousarray(noise, dtype=np.float32).reshape(B, LATENT_DIM)
        if frames_in is not None:
            frames_in = np.ascontiguousarray(frames_in, dtype=np.float32).reshape(B, LATENT_DIM)
        arr = (C.c_void_p * B)(*[s._h for s in seqs])
        lat = np.empty((B, LATENT_DIM), dtype=np.float32)
        eos = np.empty(B, dtype=np.float32)
        _check(self._lib.ptts_step(self._h, arr, B, _fptr(frames_in), _fptr(noise), lsd_steps, _fptr(lat), _fptr(eos)))
        return lat, eos

    def decode_latents(self, latents: np.ndarray, frames_per_seq: Optional[Sequence[int]] = None) -> np.ndarray:
        """LatentToMimi + MimiDecode of caller latents [N,32] (ragged batch)."""
        lat = np.ascontiguousarray(latents, dtype=np.float32).reshape(-1, LATENT_DIM)
        fps = np.asarray(frames_per_seq if frames_per_seq is not None else [lat.shape[0]], dtype=np.int32)
        assert int(fps.sum()) == lat.shape[0]
        pcm = np.empty(lat.shape[0] * SAMPLES_PER_FRAME, dtype=np.float32)
        _check(self._lib.ptts_decode_latents(self._h, _fptr(lat), fps.ctypes.data_as(C.POINTER(C.c_int32)),
                                             fps.size, _fptr(pcm)))
        return pcm

    def encode_to_latent(self, pcm: np.ndarray) -> np.ndarray:
        """MimiModel.EncodeToLatent (mimi.go:791-794): not implemented in the reference's native backend either;
        raises with ErrMimiEncoderNotImplemented's text."""
        a = np.ascontiguousarray(pcm, dtype=np.float32).ravel()
        out = np.empty((max(a.size // SAMPLES_PER_FRAME, 1), LATENT_DIM), np.float32)
        _check(self._lib.ptts_mimi_encode_to_latent(self._h, _fptr(a), C.c_int64(a.size), _fptr(out), out.shape[0]))
        return out

    def stage_times(self, reset: bool = False) -> dict:
        t = PttsStageTimes()
        _check(self._lib.ptts_stage_times_get(self._h, C.byref(t), 1 if reset else 0))
        return {n: getattr(t, n) for n, _ in PttsStageTimes._fields_}

    def debug_gemm(self, A, W, bias=None, act: int = 0, path: int = 0) -> np.ndarray:
        """C = act(A @ W.T + bias) through path 0 (auto) / 1 (CUDA cores) / 2 (tcgen05)."""
        A = np.ascontiguousarray(A, dtype=np.float32)
        W = np.ascontiguousarray(W, dtype=np.float32)
        b = None if bias is None else np.ascontiguousarray(bias, dtype=np.float32)
        out = np.empty((A.shape[0], W.shape[0]), dtype=np.float32)
        _check(self._lib.ptts_debug_gemm(self._h, _fptr(A), A.shape[0], A.shape[1], _fptr(W), W.shape[0],
                                         _fptr(b), act, path, _fptr(out)))
        return out

    def debug_gemm_tma(self, A, W, bias=None, act: int = 0, taps: int = 1, want_split: bool = False):
        """TMA-fed tcgen05 GEMM (A split into bf16 planes on the device).  taps > 1: causal conv form,
        W is [N, taps*cin].  Returns C, or (C, split copy of C) when want_split."""
        A = np.ascontiguousarray(A, dtype=np.float32)
        W = np.ascontiguousarray(W, dtype=np.float32)
        b = None if bias is None else np.ascontiguousarray(bias, dtype=np.float32)
        out = np.empty((A.shape[0], W.shape[0]), dtype=np.float32)
        sp = np.empty_like(out) if want_split else None
        _check(self._lib.ptts_debug_gemm_tma(self._h, _fptr(A), A.shape[0], A.shape[1], taps, _fptr(W), W.shape[0],
                                             _fptr(b), act, _fptr(out), _fptr(sp)))
        return (out, sp) if want_split else out

    def voice_cache_stats(self) -> dict:
        """Shared voice-prefix cache: sequences admitted on shared KV pages, prefill positions saved, entries."""
        h, t, n = C.c_int64(), C.c_int64(), C.c_int32()
        _check(self._lib.ptts_voice_cache_stats(self._h, C.byref(h), C.byref(t), C.byref(n)))
        return {"hits": h.value, "shared_tokens": t.value, "entries": n.value}

    def tc_force(self, bn: int = 0, s: int = 0):
        _check(self._lib.ptts_debug_tc_force(bn, s))

    def gemm_bench(self, M: int, K: int, N: int, path: int = 0, reps: int = 20, lda: int = 0) -> float:
        """Average microseconds per launch of one GEMM shape."""
        us = C.c_float(0)
        _check(self._lib.ptts_debug_gemm_bench(self._h, M, K, N, lda, path, reps, C.byref(us)))
        return float(us.value)

    def profile(self, on: bool):
        _check(self._lib.ptts_profile_enable(self._h, 1 if on else 0))

    def profile_report(self) -> dict:
        import json
        buf = C.create_string_buffer(1 << 16)
        n = self._lib.ptts_profile_report(self._h, buf, len(buf))
        if n < 0:
            _check(n)
        return json.loads(buf.value.decode())

    def debug_fetch(self, name: str) -> np.ndarray:
        n = self._lib.ptts_debug_fetch(self._h, name.encode(), None, 0)
        if n < 0:
            raise KeyError(name)
        out = np.empty(n, dtype=np.float32)
        self._lib.ptts_debug_fetch(self._h, name.encode(), _fptr(out), n)
        return out


class NativeCudaPool(NativeCudaRuntime):
    """One process, several GPUs: a `tts.Runtime` whose GenerateAudio calls are dispatched
    to the least-loaded of N engines (one per device, weights replicated).  submit / wait / generate /
    cancel are those of NativeCudaRuntime; `engine(i)` borrows one engine for stage times and stats."""

    def __init__(self, checkpoint: Optional[str] = None, *, devices: Optional[Sequence[int]] = None,
                 max_batch: int = 256, max_seq_tokens: int = 1024, max_frames: int = 512, mimi_group_frames: int = 0,
                 kv_pages: int = 0, tensors: Optional[dict] = None, flags: int = 0):
        self._lib = lib()
        self._h = C.c_void_p()
        cfg = PttsConfig(device=0, max_batch=max_batch, max_seq_tokens=max_seq_tokens, max_frames=max_frames,
                         mimi_group_frames=mimi_group_frames, kv_pages=kv_pages, flags=flags)
        devs = None if devices is None else (C.c_int32 * len(devices))(*devices)
        _check(self._lib.ptts_pool_create(C.byref(cfg), devs, 0 if devices is None else len(devices), C.byref(self._h)))
        try:
            if checkpoint is not None:
                _check(self._lib.ptts_pool_load_safetensors(self._h, checkpoint.encode()))
            for k, a in (tensors or {}).items():
                a = np.ascontiguousarray(a, dtype=np.float32)
                shape = (C.c_int64 * a.ndim)(*a.shape)
                _check(self._lib.ptts_pool_load_tensor(self._h, k.encode(), shape, a.ndim, 0, a.ctypes.data_as(C.c_void_p)))
            _check(self._lib.ptts_pool_finalize_weights(self._h))
            info = PttsModelInfo()
            _check(self._lib.ptts_model_info_get(self._engine_handle(0), C.byref(info)))
            self.info = info
        except Exception:
            self.close()
            raise

    @property
    def size(self) -> int:
        return self._lib.ptts_pool_size(self._h)

    def _engine_handle(self, i: int):
        h = C.c_void_p()
        _check(self._lib.ptts_pool_engine(self._h, i, C.byref(h)))
        return h

    def engine(self, i: int) -> NativeCudaRuntime:
        """Borrowed view of engine i (stage_times, voice_cache_stats, profile); do not close it."""
        v = NativeCudaRuntime.__new__(NativeCudaRuntime)
        v._lib, v._h, v.info = self._lib, self._engine_handle(i), self.info
        v.close = lambda: None
        return v

    def close(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            self._lib.ptts_pool_destroy(h)

    def load(self) -> dict:
        n = self.size
        cost, req, disp = (C.c_int64 * n)(), (C.c_int32 * n)(), (C.c_int64 * n)()
        self._lib.ptts_pool_load(self._h, cost, req, disp, n)
        return {"inflight_cost": list(cost), "inflight_requests": list(req), "dispatched": list(disp)}

    def stage_times(self, reset: bool = False) -> dict:
        """Sum over the engines (they run concurrently: the ms figures are device time, not wall time)."""
        tot = None
        for i in range(self.size):
            t = self.engine(i).stage_times(reset)
            tot = t if tot is None else {k: tot[k] + t[k] for k in t}
        return tot

    def _submit(self, *a):
        return self._lib.ptts_pool_submit(self._h, *a, None)

    def _wait(self, ticket, res):
        return self._lib.ptts_pool_wait(self._h, ticket, res)

    def _ticket_error(self, ticket, buf, cap):
        return self._lib.ptts_pool_ticket_error(self._h, ticket, buf, cap)

    def _ticket_latents(self, ticket, lat, eos, cap):
        return self._lib.ptts_pool_ticket_latents(self._h, ticket, lat, eos, cap)

    def _ticket_release(self, ticket):
        return self._lib.ptts_pool_ticket_release(self._h, ticket)

    def _ticket_read(self, ticket, have, timeout, ready, done):
        return self._lib.ptts_pool_ticket_read(self._h, ticket, have, timeout, ready, done)

    def _cancel(self, ticket):
        return self._lib.ptts_pool_cancel(self._h, ticket)


class Sequence_:
    """FlowLMState mirror (internal/native/flow_lm.go:45-49)."""

    def __init__(self, rt: NativeCudaRuntime, h):
        self._rt, self._h = rt, h

    @property
    def frames(self) -> int:
        return self._rt._lib.ptts_seq_frames(self._h)

    @property
    def kv_len(self) -> int:
        return self._rt._lib.ptts_seq_kv_len(self._h)

    def read_kv(self, layer: int, which: int) -> np.ndarray:
        info = self._rt.info
        out = np.empty((info.n_heads, self.kv_len, info.head_dim), dtype=np.float32)
        _check(self._rt._lib.ptts_seq_read_kv(self._rt._h, self._h, layer, which, _fptr(out)))
        return out

    def decode(self, first_frame: int = 0, n_frames: Optional[int] = None) -> np.ndarray:
        n = self.frames - first_frame if n_frames is None else n_frames
        pcm = np.empty(n * SAMPLES_PER_FRAME, dtype=np.float32)
        _check(self._rt._lib.ptts_decode(self._rt._h, self._h, first_frame, n, _fptr(pcm)))
        return pcm

    def close(self):
        h, self._h = self._h, None
        if h and self._rt._h:
            self._rt._lib.ptts_seq_close(self._rt._h, h)
