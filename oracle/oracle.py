"""ctypes binding of the CPU restatement (oracle/pockettts_oracle.c).

TEST INFRASTRUCTURE ONLY — see the header of pockettts_oracle.c.  Imported by
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs; never by the product package.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libpockettts_oracle.so")
_lib = None

fp = C.POINTER(C.c_float)
i64p = C.POINTER(C.c_int64)


def build(force: bool = False) -> str:
    src = os.path.join(HERE, "pockettts_oracle.c")
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
        subprocess.run(["make", "-C", HERE, "-B", "libpockettts_oracle.so"], check=True,
                       stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    return LIB


class GenCfg(C.Structure):
    _fields_ = [("temperature", C.c_double), ("eos_threshold", C.c_double), ("max_steps", C.c_int),
                ("estimated_max_steps", C.c_int), ("lsd_decode_steps", C.c_int), ("frames_after_eos", C.c_int),
                ("force_eos_step", C.c_int)]


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB)
        L.po_last_error.restype = C.c_char_p
        L.po_dot.restype = C.c_float
        L.po_dot.argtypes = [fp, fp, C.c_int]
        L.po_axpy.argtypes = [fp, C.c_float, fp, C.c_int]
        L.po_model_new.restype = C.c_void_p
        L.po_model_free.argtypes = [C.c_void_p]
        L.po_model_set_tensor.argtypes = [C.c_void_p, C.c_char_p, fp, i64p, C.c_int]
        L.po_model_finalize.argtypes = [C.c_void_p]
        L.po_model_dim.argtypes = [C.c_void_p, C.c_char_p]
        L.po_state_new.restype = C.c_void_p
        L.po_state_new.argtypes = [C.c_void_p]
        L.po_state_free.argtypes = [C.c_void_p]
        L.po_state_offset.restype = C.c_int64
        L.po_state_offset.argtypes = [C.c_void_p, C.c_int]
        L.po_state_capacity.argtypes = [C.c_void_p, C.c_int]
        L.po_state_read_kv.argtypes = [C.c_void_p, C.c_int, C.c_int, fp]
        L.po_state_set_voice_layer.argtypes = [C.c_void_p, C.c_int, fp, i64p, C.c_int, C.c_float]
        L.po_text_embeddings.argtypes = [C.c_void_p, i64p, C.c_int, fp]
        L.po_prompt.argtypes = [C.c_void_p, fp, C.c_int]
        L.po_step.argtypes = [C.c_void_p, fp, fp, C.c_int, fp, fp, fp]
        L.po_flow_direction.argtypes = [C.c_void_p, fp, C.c_float, C.c_float, fp, fp]
        L.po_lsd_decode.argtypes = [C.c_void_p, fp, fp, C.c_int, fp]
        L.po_latent_to_mimi.argtypes = [C.c_void_p, fp, C.c_int, fp]
        L.po_latent_to_mimi_unfused.argtypes = [C.c_void_p, fp, C.c_int, fp]
        L.po_mimi_decode.argtypes = [C.c_void_p, fp, C.c_int, fp, fp, fp, fp]
        L.po_generate_audio.argtypes = [C.c_void_p, C.c_void_p, i64p, C.c_int, fp, C.c_int, C.POINTER(GenCfg),
                                        fp, C.c_int, fp, C.c_long, C.POINTER(C.c_int), fp, fp]
        L.po_estimate_max_frames.argtypes = [C.c_int, C.c_double]
        L.po_softmax_lastdim.argtypes = [fp, C.c_int, C.c_int]
        L.po_layernorm.argtypes = [fp, fp, C.c_int, C.c_int, fp, fp, C.c_float]
        L.po_linear.argtypes = [fp, C.c_int, C.c_int, fp, fp, C.c_int, fp]
        L.po_build_rope.argtypes = [C.c_int, C.c_int, C.c_double, fp, fp]
        L.po_rope.argtypes = [fp, C.c_int, C.c_int, C.c_int, fp, fp, C.c_int]
        L.po_attention_positions.argtypes = [fp, fp, fp] + [C.c_int] * 6 + [i64p, i64p, C.c_int64, fp]
        L.po_attention_causal.argtypes = [fp, fp, fp] + [C.c_int] * 7 + [C.c_int64, fp]
        L.po_conv1d.argtypes = [fp, C.c_int, C.c_int, C.c_int, fp, C.c_int, C.c_int, C.c_int, fp] + [C.c_int] * 5 + [fp]
        L.po_convtranspose1d.argtypes = [fp, C.c_int, C.c_int, C.c_int, fp, C.c_int, C.c_int, fp] + [C.c_int] * 6 + [fp]
        L.po_repack_convtr.argtypes = [fp, C.c_int, C.c_int, C.c_int, fp]
        L.po_rmsnorm_alpha.argtypes = [fp, C.c_int, C.c_int, fp, C.c_float]
        L.po_mlp_silu.argtypes = [fp, C.c_int, C.c_int, fp, fp, C.c_int, fp, fp, C.c_int, fp]
        for n in ("po_silu", "po_gelu", "po_elu"):
            getattr(L, n).argtypes = [fp, C.c_int]
        L.po_set_workers.argtypes = [C.c_int]
        L.po_split_voice_kv.argtypes = [fp, C.c_int, C.c_int, C.c_int, fp, fp]
        _lib = L
    return _lib


class OracleError(RuntimeError):
    pass


def f32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float32)


def P(a):
    return None if a is None else a.ctypes.data_as(fp)


def PI(a):
    return a.ctypes.data_as(i64p)


def _chk(rc):
    if rc < 0:
        raise OracleError(lib().po_last_error().decode())
    return rc


def set_workers(n: int):
    lib().po_set_workers(int(n))


# ---- op-level wrappers (used by the KAT tests) ------------------------------
def dot(a, b) -> float:
    a, b = f32(a), f32(b)
    return float(lib().po_dot(P(a), P(b), a.size))


def axpy(dst, alpha, src):
    dst, src = f32(dst).copy(), f32(src)
    lib().po_axpy(P(dst), alpha, P(src), min(dst.size, src.size))
    return dst


def softmax(x):
    x = f32(x).copy()
    _chk(lib().po_softmax_lastdim(P(x), x.size // x.shape[-1], x.shape[-1]))
    return x


def layernorm(x, w=None, b=None, eps=1e-5):
    x = f32(x)
    out = np.empty_like(x)
    w = None if w is None else f32(w)
    b = None if b is None else f32(b)
    lib().po_layernorm(P(x), P(out), x.size // x.shape[-1], x.shape[-1], P(w), P(b), eps)
    return out


def linear(x, w, b=None):
    x, w = f32(x), f32(w)
    b = None if b is None else f32(b)
    rows = x.size // x.shape[-1]
    y = np.empty(x.shape[:-1] + (w.shape[0],), dtype=np.float32)
    lib().po_linear(P(x), rows, x.shape[-1], P(w), P(b), w.shape[0], P(y))
    return y


def build_rope(max_seq, head_dim, max_period=10000.0):
    c = np.empty((max_seq, head_dim // 2), np.float32)
    s = np.empty_like(c)
    lib().po_build_rope(max_seq, head_dim, max_period, P(c), P(s))
    return c, s


def rope(x, cos, sin, pos=0):
    """x [..., seq, dim]."""
    x = f32(x).copy()
    seq, dim = x.shape[-2], x.shape[-1]
    cos, sin = f32(cos), f32(sin)
    if cos.shape[0] < pos + seq:
        raise OracleError("ops: rope cos/sin sequence length too small")
    lib().po_rope(P(x), x.size // (seq * dim), seq, dim, P(cos), P(sin), pos)
    return x


def attention_positions(q, k, v, pos_q, pos_k, context=-1):
    q, k, v = f32(q), f32(k), f32(v)
    B, H, tq, d = q.shape
    tk, dv = k.shape[2], v.shape[3]
    pq = np.ascontiguousarray(pos_q, dtype=np.int64)
    pk = np.ascontiguousarray(pos_k, dtype=np.int64)
    out = np.empty((B, H, tq, dv), np.float32)
    _chk(lib().po_attention_positions(P(q), P(k), P(v), B, H, tq, tk, d, dv, PI(pq), PI(pk), context, P(out)))
    return out


def attention_causal(q, k, v, causal=True, offset=0):
    q, k, v = f32(q), f32(k), f32(v)
    B, H, tq, d = q.shape
    tk, dv = k.shape[2], v.shape[3]
    out = np.empty((B, H, tq, dv), np.float32)
    _chk(lib().po_attention_causal(P(q), P(k), P(v), B, H, tq, tk, d, dv, int(causal), offset, P(out)))
    return out


def conv1d(x, w, b=None, stride=1, lpad=0, rpad=0, dilation=1, groups=1):
    x, w = f32(x), f32(w)
    b = None if b is None else f32(b)
    B, Cin, L = x.shape
    Cout, kCin, k = w.shape
    out_len = (L + lpad + rpad - dilation * (k - 1) - 1) // stride + 1
    out = np.empty((B, Cout, max(out_len, 1)), np.float32)
    got = _chk(lib().po_conv1d(P(x), B, Cin, L, P(w), Cout, kCin, k, P(b), stride, lpad, rpad, dilation, groups, P(out)))
    assert got == out_len
    return out


def convtranspose1d(x, w, b=None, stride=1, padding=0, output_padding=0, dilation=1, groups=1, right_trim=0):
    x, w = f32(x), f32(w)
    b = None if b is None else f32(b)
    B, Cin, L = x.shape
    _, opg, k = w.shape
    out_len = (L - 1) * stride - 2 * padding + dilation * (k - 1) + output_padding + 1 - right_trim
    out = np.empty((B, opg * groups, max(out_len, 1)), np.float32)
    got = _chk(lib().po_convtranspose1d(P(x), B, Cin, L, P(w), opg, k, P(b), stride, padding, output_padding,
                                        dilation, groups, right_trim, P(out)))
    assert got == out_len
    return out


def repack_convtr(w):
    w = f32(w)
    Cin, Cout, k = w.shape
    out = np.empty((k, Cout, Cin), np.float32)
    lib().po_repack_convtr(P(w), Cin, Cout, k, P(out))
    return out


def mlp_silu(x, w1, b1, w2, b2):
    x, w1, w2 = f32(x), f32(w1), f32(w2)
    b1 = None if b1 is None else f32(b1)
    b2 = None if b2 is None else f32(b2)
    rows = x.size // x.shape[-1]
    y = np.empty(x.shape[:-1] + (w2.shape[0],), np.float32)
    lib().po_mlp_silu(P(x), rows, x.shape[-1], P(w1), P(b1), w1.shape[0], P(w2), P(b2), w2.shape[0], P(y))
    return y


def _unary(name, x):
    x = f32(x).copy()
    getattr(lib(), name)(P(x), x.size)
    return x


def silu(x):
    return _unary("po_silu", x)


def gelu(x):
    return _unary("po_gelu", x)


def elu(x):
    return _unary("po_elu", x)


def rmsnorm_alpha(x, alpha, eps=1e-5):
    x = f32(x).copy()
    alpha = f32(alpha)
    lib().po_rmsnorm_alpha(P(x), x.size // x.shape[-1], x.shape[-1], P(alpha), eps)
    return x


def split_voice_kv(cache):
    """cache [2,1,T,H,D] -> (K, V) each [H,T,D]."""
    cache = f32(cache)
    _, _, T, H, D = cache.shape
    k = np.empty((H, T, D), np.float32)
    v = np.empty_like(k)
    lib().po_split_voice_kv(P(cache), T, H, D, P(k), P(v))
    return k, v


def estimate_max_frames(n, rate=12.5):
    return lib().po_estimate_max_frames(n, rate)


# ---- model-level ------------------------------------------------------------
class OracleModel:
    """native.Model restated (internal/native/model.go:25-167)."""

    def __init__(self, tensors: dict):
        L = lib()
        self._m = C.c_void_p(L.po_model_new())
        for k, a in tensors.items():
            if a.dtype != np.float32:
                continue
            a = f32(a)
            shape = np.asarray(a.shape, dtype=np.int64)
            _chk(L.po_model_set_tensor(self._m, k.encode(), P(a), PI(shape), a.ndim))
        _chk(L.po_model_finalize(self._m))
        self.D = self.dim("D")
        self.ldim = self.dim("ldim")
        self.Md = self.dim("Md")
        self.L = self.dim("L")
        self.H = self.dim("H")
        self.Dh = self.dim("Dh")

    def dim(self, what):
        return lib().po_model_dim(self._m, what.encode())

    def close(self):
        if self._m:
            lib().po_model_free(self._m)
            self._m = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def text_embeddings(self, tokens):
        t = np.ascontiguousarray(tokens, dtype=np.int64)
        out = np.empty((t.size, self.D), np.float32)
        _chk(lib().po_text_embeddings(self._m, PI(t), t.size, P(out)))
        return out

    def new_state(self):
        return OracleState(self)

    def flow_direction(self, c, s, t, x):
        c, x = f32(c), f32(x)
        out = np.empty(self.ldim, np.float32)
        _chk(lib().po_flow_direction(self._m, P(c), s, t, P(x), P(out)))
        return out

    def lsd_decode(self, c, x0, steps):
        c, x0 = f32(c), f32(x0)
        out = np.empty(self.ldim, np.float32)
        _chk(lib().po_lsd_decode(self._m, P(c), P(x0), steps, P(out)))
        return out

    def latent_to_mimi(self, latent, unfused=False):
        latent = f32(latent).reshape(-1, self.ldim)
        T = latent.shape[0]
        out = np.empty((self.Md, T), np.float32)
        fn = lib().po_latent_to_mimi_unfused if unfused else lib().po_latent_to_mimi
        _chk(fn(self._m, P(latent), T, P(out)))
        return out

    def mimi_decode(self, mimi_latent, stages=False):
        ml = f32(mimi_latent)
        T = ml.shape[1]
        pcm = np.empty(1920 * T, np.float32)
        if stages:
            s0 = np.empty((self.Md, 16 * T), np.float32)
            s1 = np.empty_like(s0)
            s2 = np.empty_like(s0)
            _chk(lib().po_mimi_decode(self._m, P(ml), T, P(pcm), P(s0), P(s1), P(s2)))
            return pcm, (s0, s1, s2)
        _chk(lib().po_mimi_decode(self._m, P(ml), T, P(pcm), None, None, None))
        return pcm

    def decode_latents(self, latents):
        return self.mimi_decode(self.latent_to_mimi(latents))

    def generate_audio(self, tokens, *, noise, max_steps=0, estimated_max_steps=0, lsd_decode_steps=1,
                       frames_after_eos=3, eos_threshold=-4.0, temperature=0.7, force_eos_step=-1,
                       voice_embedding=None, state=None, want_pcm=True):
        """GenerateAudio restated (internal/tts/runtime_native_safetensors.go:52-238).
        Returns (pcm, latents [n,32], eos_logits [n])."""
        t = np.ascontiguousarray(tokens, dtype=np.int64)
        noise = f32(noise).reshape(-1, self.ldim)
        cfg = GenCfg(temperature, eos_threshold, max_steps, estimated_max_steps, lsd_decode_steps,
                     frames_after_eos, force_eos_step)
        cap = noise.shape[0]
        pcm = np.empty(cap * 1920, np.float32) if want_pcm else None
        lat = np.empty((cap, self.ldim), np.float32)
        eos = np.empty(cap, np.float32)
        nf = C.c_int(0)
        ve, tv = None, 0
        if voice_embedding is not None:
            ve = f32(voice_embedding).reshape(-1, self.D)
            tv = ve.shape[0]
        _chk(lib().po_generate_audio(self._m, state._s if state is not None else None, PI(t), t.size, P(ve), tv,
                                     C.byref(cfg), P(noise), noise.shape[0], P(pcm), 0 if pcm is None else pcm.size,
                                     C.byref(nf), P(lat), P(eos)))
        n = nf.value
        return (pcm[: n * 1920].copy() if pcm is not None else None), lat[:n].copy(), eos[:n].copy()


class OracleState:
    """FlowLMState restated (flow_lm.go:45-49, flow_transformer.go:26-108)."""

    def __init__(self, model: OracleModel):
        self.m = model
        self._s = C.c_void_p(lib().po_state_new(model._m))

    def close(self):
        if self._s:
            lib().po_state_free(self._s)
            self._s = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_voice_layer(self, layer, cache, offset):
        cache = f32(cache)
        shape = np.asarray(cache.shape, dtype=np.int64)
        _chk(lib().po_state_set_voice_layer(self._s, layer, P(cache), PI(shape), cache.ndim, float(offset)))

    def offset(self, layer=0):
        return int(lib().po_state_offset(self._s, layer))

    def capacity(self, layer=0):
        return int(lib().po_state_capacity(self._s, layer))

    def read_kv(self, layer, which):
        out = np.empty((self.m.H, self.offset(layer), self.m.Dh), np.float32)
        lib().po_state_read_kv(self._s, layer, which, P(out))
        return out

    def prompt(self, emb):
        emb = f32(emb).reshape(-1, self.m.D)
        _chk(lib().po_prompt(self._s, P(emb), emb.shape[0]))

    def step(self, frame_in, noise, lsd_steps=1):
        """Returns (latent [32], eos_logit, last_hidden [D])."""
        frame_in, noise = f32(frame_in), f32(noise)
        lat = np.empty(self.m.ldim, np.float32)
        eos = C.c_float(0)
        last = np.empty(self.m.D, np.float32)
        _chk(lib().po_step(self._s, P(frame_in), P(noise), lsd_steps, P(lat), C.byref(eos), P(last)))
        return lat, float(eos.value), last
