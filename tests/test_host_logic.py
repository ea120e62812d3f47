"""CPU tests of the host side: the C-ABI library loads and exports every symbol
the header declares (no compute calls without a GPU), request-config resolution,
synthetic checkpoint / safetensors I/O, workload generator."""
import ctypes
import importlib
import os
import re

import numpy as np
import pytest

import __graft_entry__ as ge


def test_library_builds_and_exports_every_declared_symbol(pkg):
    pkg.build_library()
    header = open(os.path.join(ge.ROOT, "include", "pockettts_cuda.h")).read()
    declared = set(re.findall(r"\b(ptts_[a-z_0-9]+)\s*\(", header))
    declared -= {"ptts_engine", "ptts_seq"}
    assert len(declared) >= 25
    lib = ctypes.CDLL(pkg.LIB_PATH)
    for sym in sorted(declared):
        assert getattr(lib, sym) is not None, sym
    assert declared == set(pkg.EXPORTS)


def test_library_has_sm100a_code_only(pkg):
    import subprocess
    out = subprocess.run(["cuobjdump", "-lelf", pkg.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    assert not re.search(r"sm_(?!100a)\d+", out)


def test_no_gpu_means_loud_failure_not_fallback(pkg):
    """The product path must fail loudly when it cannot run on a GPU."""
    if pkg.lib().ptts_device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(pkg.NativeCudaError, match="runtime unavailable"):
        pkg.NativeCudaRuntime(tensors=None)


def test_product_never_imports_oracle():
    for root, _, files in os.walk(ge.PKG_DIR):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".go")):
                src = open(os.path.join(root, f), errors="ignore").read()
                assert "pockettts_oracle" not in src and "po_generate_audio" not in src, f
                assert not re.search(r"^\s*(from|import)\s+oracle", src, re.M), f


def test_max_step_resolution(pkg):
    # runtime_native_safetensors.go:61-67 + prepare.go:38-48
    C = pkg.RuntimeGenerateConfig
    assert pkg.resolve_max_steps(C(max_steps=7, estimated_max_steps=9), 11) == 7
    assert pkg.resolve_max_steps(C(estimated_max_steps=9), 11) == 9
    assert pkg.resolve_max_steps(C(), 11) == 71
    assert pkg.estimate_max_frames(50) == 234
    assert pkg.estimate_max_frames(-1) == 25


def test_noise_matches_reference_contract(pkg):
    # flow_lm.go:386-408: sigma = sqrt(max(T,0)); negative temperature -> zeros
    a = pkg.make_noise(5, 0.7, 3)
    b = pkg.make_noise(5, 0.7, 3)
    assert a.shape == (5, 32) and a.dtype == np.float32
    np.testing.assert_array_equal(a, b)
    assert np.all(pkg.make_noise(2, -0.5, 1) == 0)
    assert abs(float(pkg.make_noise(4000, 0.7, 9).std()) - np.sqrt(0.7)) < 0.01


def test_voice_embedding_shape_normalisation(pkg):
    # internal/safetensors/reader.go:219-230: [T,D] -> [1,T,D]; other ranks rejected
    v = np.zeros((5, 1024), np.float32)
    assert pkg.VoiceEmbedding(v).rows().shape == (5, 1024)
    assert pkg.VoiceEmbedding(v[None]).rows().shape == (5, 1024)
    with pytest.raises(ValueError):
        pkg.VoiceEmbedding(np.zeros(7, np.float32)).rows()


def test_checkpoint_has_every_key_the_loader_reads(ck):
    t = ck.make_checkpoint(ck.ArchConfig.tiny(), seed=1)
    need = ["flow_lm.conditioner.embed.weight", "flow_lm.emb_std", "flow_lm.emb_mean", "flow_lm.bos_emb",
            "flow_lm.input_linear.weight", "flow_lm.out_norm.weight", "flow_lm.out_eos.bias",
            "flow_lm.transformer.layers.1.self_attn.in_proj.weight", "flow_lm.transformer.layers.0.linear2.weight",
            "flow_lm.flow_net.time_embed.1.mlp.3.alpha", "flow_lm.flow_net.res_blocks.1.adaLN_modulation.1.bias",
            "flow_lm.flow_net.final_layer.linear.weight", "mimi.quantizer.output_proj.weight",
            "mimi.upsample.convtr.convtr.weight",
            "mimi.decoder_transformer.transformer.layers.0.layer_scale_2.scale",
            "mimi.decoder.model.0.conv.weight", "mimi.decoder.model.2.convtr.bias",
            "mimi.decoder.model.9.block.3.conv.weight", "mimi.decoder.model.11.conv.bias"]
    for k in need:
        assert k in t, k
    full = ck.ArchConfig.full()
    assert (full.d_model, full.n_layers, full.ffn_dim, full.flow_dim, full.flow_depth) == (1024, 6, 4096, 512, 6)


def test_full_architecture_param_count_matches_survey(ck):
    """P_AR = 85,316,897 with the inferred shapes (computed from shapes only)."""
    c = ck.ArchConfig.full()
    D, F, C, L = c.d_model, c.ffn_dim, c.flow_dim, c.n_layers
    transformer = L * (4 * D * D + 2 * D * F + 4 * D)
    io = 32 * D + 2 * D + (D + 1)
    flow = (32 * C + C) + (D * C + C) + 2 * (128 + 256 * C + C + C * C + C + C) + \
        6 * (2 * C + 2 * (C * C + C) + 3 * C * C + 3 * C) + (32 * C + 32) + (2 * C * C + 2 * C)
    assert transformer + io + flow == 85_316_897


def test_safetensors_roundtrip_and_bf16_widening(ck, tmp_path):
    t = ck.make_checkpoint(ck.ArchConfig.tiny(), seed=2)
    p = str(tmp_path / "a.safetensors")
    ck.write_safetensors(p, t)
    r = ck.read_safetensors(p)
    assert set(r) == set(t)
    for k in t:
        np.testing.assert_array_equal(r[k], t[k])
    bf = ck.to_bf16_exact(t)
    ck.write_safetensors(p, bf, bf16=True)
    r = ck.read_safetensors(p)
    for k in t:
        np.testing.assert_array_equal(r[k], bf[k])  # store.go:339-395: BF16 widens exactly
        assert np.abs(bf[k] - t[k]).max() <= np.abs(t[k]).max() * 2 ** -8 + 1e-30


def test_workload_follows_reference_chunking_rules():
    ge.package()
    wl = importlib.import_module("pockettts_b200.workload")
    chunks = wl.make_workload(200, 5, 60, seed=7)
    assert len({c.utt for c in chunks}) == 200
    for c in chunks:
        n = c.tokens.size
        assert n <= 50 + 8
        assert c.max_steps == int(np.ceil((n / 3 + 2) * 12.5))      # prepare.go:38-48
        assert c.frames_after_eos == (5 if c.n_words <= 4 else 3)     # prepare.go:53-59
        assert 0 < c.expected_frames <= c.max_steps
        assert c.noise().shape == (c.max_steps, 32)
    hw = wl.hello_world()
    assert hw.tokens.tolist() == [260] * 8 + [2994, 578, 263] and hw.max_steps == 71 and hw.frames_after_eos == 5
    c2 = wl.paragraph_200_words()
    assert 5 <= len(c2) <= 7 and sum(c.n_words for c in c2) == 200
    # determinism of the shard: same seed, same chunks
    again = wl.make_workload(200, 5, 60, seed=7)
    assert all(np.array_equal(a.tokens, b.tokens) for a, b in zip(chunks, again))


def test_bench_reference_arm_prints_exactly_one_json_line():
    """bench.py contract: stdout carries ONE JSON line whatever the libraries print (NCCL's banner
    goes to fd 1); checked on the CPU-only reference arm with a tiny sample."""
    import json, os, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1",
                        "--warmup", "0", "--arch", "small", "--cpu-sample-frames", "2"],
                       capture_output=True, text=True, timeout=600, cwd=root)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout[:500]
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "audio-s/s" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["e2e"]["h2d_bytes_per_step"] == 0
