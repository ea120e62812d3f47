"""Whole-path checks at the FULL PocketTTS architecture (BASELINE.json's sizes: 6 x 1024 FlowLM,
flow depth 6, Mimi 2 x 512 + SEANet), where the CPU oracle is too slow to be the checker for
whole utterances: size-independent properties of the path plus one short oracle comparison."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def snr_db(ref, got):
    ref = ref.astype(np.float64)
    err = ref - got.astype(np.float64)
    return 10 * np.log10(float(np.sum(ref ** 2)) / max(float(np.sum(err ** 2)), 1e-30))


@pytest.fixture(scope="module")
def full_tensors(ck):
    return ck.make_checkpoint(ck.ArchConfig.full())


@pytest.fixture(scope="module")
def full_rt(pkg, full_tensors):
    rt = pkg.NativeCudaRuntime(tensors=full_tensors, max_batch=64, max_seq_tokens=512, max_frames=128,
                               mimi_group_frames=1024)
    yield rt
    rt.close()


def _cfg(pkg, steps, force, seed, **kw):
    return pkg.RuntimeGenerateConfig(max_steps=steps, eos_threshold=1e30, force_eos_step=force, frames_after_eos=2,
                                     noise=pkg.make_noise(steps, 0.7, seed), **kw)


def _tokens(n, seed):
    return np.random.default_rng(seed).integers(0, 4000, n).astype(np.int64)


def test_model_constants_match_survey(full_rt):
    """P_AR = 85,316,897 parameters; 270.7 M MAC per Mimi frame of which <= 8.2 M are the
    window attention (the engine counts the 262.5 M GEMM / conv MACs), recomputed from the loaded shapes."""
    info = full_rt.info
    assert info.params_ar == 85316897
    assert info.macs_mimi_frame == 262529024
    assert (info.d_model, info.n_layers, info.ffn_dim, info.flow_dim, info.mimi_layers) == (1024, 6, 4096, 512, 2)


def test_short_utterance_matches_oracle_at_full_size(full_rt, full_tensors, orc, pkg):
    """One short request end to end against the CPU restatement (latents 1e-3, waveform >= 40 dB)."""
    om = orc.OracleModel(full_tensors)
    toks = _tokens(5, 1)
    cfg = _cfg(pkg, 4, 1, 2)
    got = full_rt.generate(toks, cfg, want_latents=True)
    pcm, lat, _ = om.generate_audio(toks, noise=cfg.noise, max_steps=4, eos_threshold=1e30, force_eos_step=1,
                                    frames_after_eos=2)
    assert got.n_frames == lat.shape[0] == 4
    assert np.all(np.abs(got.latents - lat) <= 2e-4 + 1e-3 * np.abs(lat)), np.abs(got.latents - lat).max()
    assert snr_db(pcm, got.pcm) >= 40.0


def test_generation_is_deterministic(full_rt, pkg):
    """Same tokens, voice and noise -> bit-identical latents and PCM (no atomics on the data path;
    the batch-1 frame is the fused cooperative kernel)."""
    toks = _tokens(30, 3)
    a = full_rt.generate(toks, _cfg(pkg, 40, 30, 4), want_latents=True)
    b = full_rt.generate(toks, _cfg(pkg, 40, 30, 4), want_latents=True)
    assert a.n_frames == b.n_frames == 33
    np.testing.assert_array_equal(a.latents, b.latents)
    np.testing.assert_array_equal(a.pcm, b.pcm)


def test_batching_does_not_change_a_request(full_rt, pkg):
    """A request generated alone (fused B=1 kernel, GEMV) and inside a batch of 40 (tcgen05 GEMMs,
    different tile / split choices) agrees within the latent tolerance after 40 fed-back steps."""
    toks = [_tokens(12 + i % 30, 100 + i) for i in range(40)]
    cfgs = [_cfg(pkg, 45, 36 + i % 5, 200 + i) for i in range(40)]
    alone = full_rt.generate(toks[7], cfgs[7], want_latents=True)
    pend = [full_rt.submit(t, c) for t, c in zip(toks, cfgs)]
    outs = [full_rt.wait(p, want_latents=True) for p in pend]
    got = outs[7]
    assert got.n_frames == alone.n_frames
    assert np.all(np.abs(got.latents - alone.latents) <= 2e-4 + 1e-3 * np.abs(alone.latents)), \
        np.abs(got.latents - alone.latents).max()
    assert snr_db(alone.pcm, got.pcm) >= 40.0
    assert all(np.isfinite(o.pcm).all() and np.abs(o.latents).max() < 10 for o in outs)


def test_streaming_window_equals_whole_decode_at_full_size(full_rt, pkg):
    """f2 at the real receptive field (2 layers x 249 rows + conv context = 33 frames)."""
    s = full_rt.seq_open(_tokens(20, 5), reserve_frames=100)
    noise = pkg.make_noise(90, 0.7, 6)
    for i in range(90):
        full_rt.step([s], noise[i])
    whole = s.decode(0, 90)
    for first, n in ((33, 1), (60, 2), (89, 1)):
        part = s.decode(first, n)
        ref = whole[first * 1920:(first + n) * 1920]
        assert snr_db(ref, part) >= 60.0, (first, snr_db(ref, part))
    s.close()


def test_pcm_is_prefix_stable(full_rt, pkg):
    """Causality of the whole path: the first k frames of a longer generation equal a generation that
    stops after k frames (same tokens and noise rows)."""
    toks = _tokens(25, 7)
    noise = pkg.make_noise(30, 0.7, 8)
    long = full_rt.generate(toks, pkg.RuntimeGenerateConfig(max_steps=30, eos_threshold=1e30, force_eos_step=-1,
                                                            frames_after_eos=2, noise=noise), want_latents=True)
    short = full_rt.generate(toks, pkg.RuntimeGenerateConfig(max_steps=12, eos_threshold=1e30, force_eos_step=-1,
                                                             frames_after_eos=2, noise=noise[:12]), want_latents=True)
    assert long.n_frames == 30 and short.n_frames == 12
    np.testing.assert_array_equal(long.latents[:12], short.latents)
    assert snr_db(long.pcm[: 12 * 1920], short.pcm) >= 60.0


def test_fused_frame_kernel_variants_agree(pkg, full_tensors, monkeypatch):
    """The batch-1 frame kernel exists twice: flag-in-data exchange (PTTS_FUSED_LL=1) and grid
    barriers (PTTS_FUSED_LL=0).  Both compute every dot product in the same order; only the LayerNorm
    partial sums are grouped differently, so 24 fed-back frames (1 and 2 LSD steps) agree far inside
    the latent tolerance, and the flag-in-data kernel is deterministic across engines."""
    toks = _tokens(22, 11)

    def run(lsd):
        rt = pkg.NativeCudaRuntime(tensors=full_tensors, max_batch=4, max_seq_tokens=256, max_frames=64,
                                   mimi_group_frames=64)
        try:
            cfg = _cfg(pkg, 24, 20, 12, lsd_decode_steps=lsd)
            return rt.generate(toks, cfg, want_latents=True)
        finally:
            rt.close()

    for lsd in (1, 2):
        monkeypatch.setenv("PTTS_FUSED_LL", "1")
        a = run(lsd)
        a2 = run(lsd)
        monkeypatch.setenv("PTTS_FUSED_LL", "0")
        b = run(lsd)
        assert a.n_frames == b.n_frames == 23
        np.testing.assert_array_equal(a.latents, a2.latents)
        assert np.all(np.abs(a.latents - b.latents) <= 1e-5 + 1e-4 * np.abs(b.latents)), \
            (lsd, np.abs(a.latents - b.latents).max())
        assert snr_db(b.pcm, a.pcm) >= 60.0


def test_weight_ring_frame_kernel_is_bit_identical(pkg, full_tensors, monkeypatch):
    """PTTS_FUSED_RING=1: the frame's weights arrive through the shared-memory ring (bulk copies issued by a
    producer thread, step_ring.cu) instead of register loads; rows belong to other warps but every dot product,
    chunk order and reduction is the one of the grid-barrier kernel, so 24 fed-back frames are bit-identical
    (1 and 2 LSD steps), and the oracle tolerance holds through it."""
    toks = _tokens(22, 11)

    def run(lsd):
        rt = pkg.NativeCudaRuntime(tensors=full_tensors, max_batch=4, max_seq_tokens=256, max_frames=64,
                                   mimi_group_frames=64)
        try:
            return rt.generate(toks, _cfg(pkg, 24, 20, 12, lsd_decode_steps=lsd), want_latents=True)
        finally:
            rt.close()

    for lsd in (1, 2):
        monkeypatch.setenv("PTTS_FUSED_LL", "0")
        monkeypatch.setenv("PTTS_FUSED_RING", "0")
        b = run(lsd)
        monkeypatch.setenv("PTTS_FUSED_RING", "1")
        a = run(lsd)
        assert a.n_frames == b.n_frames == 23
        np.testing.assert_array_equal(a.latents, b.latents)
        np.testing.assert_array_equal(a.eos_logits, b.eos_logits)
        np.testing.assert_array_equal(a.pcm, b.pcm)
