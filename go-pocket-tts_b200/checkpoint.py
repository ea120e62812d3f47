"""Seeded random-init PocketTTS-architecture checkpoints + a tiny safetensors
reader/writer.

Neither the real checkpoint (`tts_b6369a24.safetensors`) nor network access
exists offline, so benchmarks and parity tests use random weights of the exact
architecture and KEY NAMES the reference loader reads (SURVEY.md §8c;
internal/native/{flow_lm.go:51-119, flow_transformer.go:110-164,
flow_net.go:18-305, mimi.go:546-637}).  Shapes marked [inferred] in SURVEY.md §8d
come from the upstream architecture; every consumer recomputes its figures
from the loaded shapes.

Init: Linear/Conv ~ N(0, 1/fan_in); residual-writing
projections x0.5; LayerNorm weight ~ 1, bias ~ 0 (small jitter so the affine
path is exercised); layer_scale 0.1; embeddings N(0,1);
freqs[j] = exp(-ln(10000) j / n).
"""
from __future__ import annotations

import json
import struct
from dataclasses import dataclass, field

import numpy as np

DEFAULT_SEED = 20260118


@dataclass
class ArchConfig:
    d_model: int = 1024
    n_layers: int = 6
    ffn_dim: int = 4096
    n_bins: int = 4000
    ldim: int = 32
    flow_dim: int = 512
    flow_depth: int = 6
    n_freqs: int = 128
    mimi_dim: int = 512
    mimi_layers: int = 2
    mimi_ffn: int = 2048
    upsample_k: int = 32
    seanet_channels: tuple = (512, 256, 128, 64)
    seanet_ratios: tuple = (6, 5, 4)
    init_k: int = 7
    res_k: int = 3
    final_k: int = 3
    with_input_bias: bool = False

    @staticmethod
    def full() -> "ArchConfig":
        return ArchConfig()

    @staticmethod
    def tiny() -> "ArchConfig":
        """CPU-test sized (oracle only: the CUDA kernels require head_dim 64)."""
        return ArchConfig(d_model=64, n_layers=2, ffn_dim=128, n_bins=50, flow_dim=32, flow_depth=2,
                          n_freqs=8, mimi_dim=32, mimi_layers=1, mimi_ffn=64, upsample_k=32,
                          seanet_channels=(32, 16, 8, 4))

    @staticmethod
    def small() -> "ArchConfig":
        """Real head_dim (64) and channel plan, fewer layers / narrower FFN: GPU
        parity cases that the CPU oracle finishes in seconds."""
        return ArchConfig(n_layers=2, ffn_dim=1024, flow_depth=2, mimi_layers=1, mimi_ffn=512)


def make_checkpoint(cfg: ArchConfig = ArchConfig(), seed: int = DEFAULT_SEED) -> dict:
    rng = np.random.default_rng(seed)
    t: dict[str, np.ndarray] = {}

    def normal(shape, std):
        return (rng.standard_normal(shape, dtype=np.float32) * np.float32(std)).astype(np.float32)

    def linear(name, out, inp, bias=False, scale=1.0):
        t[name + ".weight"] = normal((out, inp), scale / np.sqrt(inp))
        if bias:
            t[name + ".bias"] = normal((out,), 0.02)

    def lnorm(name, d):
        t[name + ".weight"] = (1.0 + normal((d,), 0.05)).astype(np.float32)
        t[name + ".bias"] = normal((d,), 0.02)

    D, F, C = cfg.d_model, cfg.ffn_dim, cfg.flow_dim
    t["flow_lm.conditioner.embed.weight"] = normal((cfg.n_bins + 1, D), 1.0)
    for i in range(cfg.n_layers):
        p = f"flow_lm.transformer.layers.{i}"
        lnorm(p + ".norm1", D)
        lnorm(p + ".norm2", D)
        linear(p + ".self_attn.in_proj", 3 * D, D)
        linear(p + ".self_attn.out_proj", D, D, scale=0.5)
        linear(p + ".linear1", F, D)
        linear(p + ".linear2", D, F, scale=0.5)
    t["flow_lm.emb_std"] = (1.0 + 0.25 * rng.random(cfg.ldim, dtype=np.float32)).astype(np.float32)
    t["flow_lm.emb_mean"] = normal((cfg.ldim,), 0.1)
    t["flow_lm.bos_emb"] = normal((cfg.ldim,), 1.0)
    linear("flow_lm.input_linear", D, cfg.ldim, bias=cfg.with_input_bias)
    lnorm("flow_lm.out_norm", D)
    linear("flow_lm.out_eos", 1, D, bias=True)
    fn = "flow_lm.flow_net"
    for i in range(2):
        p = f"{fn}.time_embed.{i}"
        j = np.arange(cfg.n_freqs, dtype=np.float64)
        t[p + ".freqs"] = np.exp(-np.log(10000.0) * j / cfg.n_freqs).astype(np.float32)
        linear(p + ".mlp.0", C, 2 * cfg.n_freqs, bias=True)
        linear(p + ".mlp.2", C, C, bias=True)
        t[p + ".mlp.3.alpha"] = (1.0 + normal((C,), 0.05)).astype(np.float32)
    linear(fn + ".cond_embed", C, D, bias=True)
    linear(fn + ".input_proj", C, cfg.ldim, bias=True)
    for i in range(cfg.flow_depth):
        p = f"{fn}.res_blocks.{i}"
        lnorm(p + ".in_ln", C)
        linear(p + ".mlp.0", C, C, bias=True)
        linear(p + ".mlp.2", C, C, bias=True, scale=0.5)
        linear(p + ".adaLN_modulation.1", 3 * C, C, bias=True, scale=0.5)
    linear(fn + ".final_layer.linear", cfg.ldim, C, bias=True, scale=0.5)
    linear(fn + ".final_layer.adaLN_modulation.1", 2 * C, C, bias=True, scale=0.5)

    Md, MF = cfg.mimi_dim, cfg.mimi_ffn
    t["mimi.quantizer.output_proj.weight"] = normal((Md, cfg.ldim, 1), 1.0 / np.sqrt(cfg.ldim))
    t["mimi.upsample.convtr.convtr.weight"] = normal((Md, 1, cfg.upsample_k), 1.0 / np.sqrt(2.0))
    for i in range(cfg.mimi_layers):
        p = f"mimi.decoder_transformer.transformer.layers.{i}"
        lnorm(p + ".norm1", Md)
        lnorm(p + ".norm2", Md)
        linear(p + ".self_attn.in_proj", 3 * Md, Md)
        linear(p + ".self_attn.out_proj", Md, Md)
        linear(p + ".linear1", MF, Md)
        linear(p + ".linear2", Md, MF)
        t[p + ".layer_scale_1.scale"] = np.full((Md,), 0.1, np.float32)
        t[p + ".layer_scale_2.scale"] = np.full((Md,), 0.1, np.float32)

    def conv(name, co, ci, k):
        t[name + ".weight"] = normal((co, ci, k), 1.0 / np.sqrt(ci * k))
        t[name + ".bias"] = normal((co,), 0.02)

    def convtr(name, ci, co, k, stride):
        # each output sample sums ci * k/stride products
        t[name + ".weight"] = normal((ci, co, k), 1.0 / np.sqrt(ci * k / stride))
        t[name + ".bias"] = normal((co,), 0.02)

    ch = cfg.seanet_channels
    conv("mimi.decoder.model.0.conv", ch[0], Md, cfg.init_k)
    for s, (up, rs) in enumerate(((2, 3), (5, 6), (8, 9))):
        r = cfg.seanet_ratios[s]
        convtr(f"mimi.decoder.model.{up}.convtr", ch[s], ch[s + 1], 2 * r, r)
        conv(f"mimi.decoder.model.{rs}.block.1.conv", ch[s + 1] // 2, ch[s + 1], cfg.res_k)
        conv(f"mimi.decoder.model.{rs}.block.3.conv", ch[s + 1], ch[s + 1] // 2, 1)
    conv("mimi.decoder.model.11.conv", 1, ch[3], cfg.final_k)
    return t


# ---------------------------------------------------------------------------
# safetensors (format: internal/safetensors/store.go:246-271, writer.go:15-83)
# ---------------------------------------------------------------------------
_DT = {"F32": np.float32, "F16": np.float16, "I64": np.int64}


def write_safetensors(path: str, tensors: dict, bf16: bool = False):
    header, off = {}, 0
    blobs = []
    for k in sorted(tensors):
        a = np.ascontiguousarray(tensors[k])
        if a.dtype == np.float32 and bf16:
            raw = (a.view(np.uint32) >> 16).astype(np.uint16).tobytes()  # truncation: exact for bf16-valued data
            dt = "BF16"
        elif a.dtype == np.float32:
            raw, dt = a.tobytes(), "F32"
        elif a.dtype == np.float16:
            raw, dt = a.tobytes(), "F16"
        elif a.dtype == np.int64:
            raw, dt = a.tobytes(), "I64"
        else:
            raise ValueError(f"unsupported dtype {a.dtype} for {k}")
        header[k] = {"dtype": dt, "shape": list(a.shape), "data_offsets": [off, off + len(raw)]}
        off += len(raw)
        blobs.append(raw)
    hj = json.dumps(header, separators=(",", ":")).encode()
    hj += b" " * ((8 - len(hj) % 8) % 8)
    with open(path, "wb") as f:
        f.write(struct.pack("<Q", len(hj)))
        f.write(hj)
        for b in blobs:
            f.write(b)


def read_safetensors(path: str) -> dict:
    """Decodes F32/F16/BF16 to float32 (internal/safetensors/store.go:339-395)."""
    with open(path, "rb") as f:
        buf = f.read()
    (hl,) = struct.unpack("<Q", buf[:8])
    header = json.loads(buf[8:8 + hl])
    data = memoryview(buf)[8 + hl:]
    out = {}
    for k, m in header.items():
        if k == "__metadata__":
            continue
        a, b = m["data_offsets"]
        raw = data[a:b]
        if m["dtype"] == "BF16":
            u = np.frombuffer(raw, dtype=np.uint16).astype(np.uint32) << 16
            arr = u.view(np.float32)
        elif m["dtype"] in _DT:
            arr = np.frombuffer(raw, dtype=_DT[m["dtype"]])
            if m["dtype"] == "F16":
                arr = arr.astype(np.float32)
        else:
            continue
        out[k] = arr.reshape(m["shape"]).copy()
    return out


def to_bf16_exact(tensors: dict) -> dict:
    """Rounds every fp32 tensor to a bf16-representable value (a checkpoint that
    is *stored* in bf16 widens to exactly these fp32 values in the reference)."""
    out = {}
    for k, a in tensors.items():
        if a.dtype == np.float32:
            u = a.view(np.uint32).astype(np.uint64)
            u = ((u + 0x7FFF + ((u >> 16) & 1)) >> 16) << 16
            out[k] = u.astype(np.uint32).view(np.float32).reshape(a.shape)
        else:
            out[k] = a
    return out
