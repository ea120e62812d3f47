"""Frame-level streaming of scheduler requests (PTTS_REQ_STREAM + ptts_ticket_read): PCM is
delivered while the AR loop is still running; the concatenated stream equals the whole-chunk decode of
the same request (causal decoder), against the oracle too."""
import numpy as np
import pytest

from conftest import allclose_ref, snr_db

pytestmark = pytest.mark.gpu


def _tokens(n, seed):
    return np.random.default_rng(seed).integers(0, 4000, size=n).astype(np.int64)


@pytest.mark.parametrize("stream_frames", [1, 3])
def test_stream_equals_whole_chunk_and_oracle(small_rt, small_oracle, pkg, stream_frames):
    toks = _tokens(13, 61)
    n = 40  # > the small model's 18-frame receptive field: deliveries with a full window
    noise = pkg.make_noise(n, 0.7, 62)
    kw = dict(max_steps=n, eos_threshold=1e30, force_eos_step=33, frames_after_eos=3, noise=noise)
    whole = small_rt.generate(toks, pkg.RuntimeGenerateConfig(**kw), want_latents=True)
    p = small_rt.submit(toks, pkg.RuntimeGenerateConfig(stream_frames=stream_frames, **kw))
    have, marks, pieces = 0, [], []
    while True:
        ready, done = small_rt.read(p, have)
        assert ready >= have
        if ready > have:
            pieces.append(p.pcm[have * 1920: ready * 1920].copy())  # valid as soon as it is announced
            marks.append(ready)
            have = ready
        if done:
            break
    got = small_rt.wait(p, want_latents=True)
    assert got.n_frames == whole.n_frames == 37 and have == 37
    assert len(marks) >= 37 // stream_frames  # really incremental, not one final delivery
    np.testing.assert_array_equal(got.latents, whole.latents)
    stream = np.concatenate(pieces)
    assert stream.shape == whole.pcm.shape
    assert snr_db(whole.pcm, stream) >= 60.0, snr_db(whole.pcm, stream)
    np.testing.assert_array_equal(stream, got.pcm)
    pcm, lat, _ = small_oracle.generate_audio(toks, **kw)
    assert allclose_ref(got.latents, lat, 2e-4, 1e-3)
    assert snr_db(pcm, stream) >= 40.0


def test_stream_generator_among_batched_requests(small_rt, pkg):
    """A streaming request batched with ordinary ones: everybody completes, the stream is contiguous."""
    others = [small_rt.submit(_tokens(5 + i, 70 + i), pkg.RuntimeGenerateConfig(max_steps=12 + i, eos_threshold=1e30, seed=i))
              for i in range(5)]
    cfg = pkg.RuntimeGenerateConfig(max_steps=25, eos_threshold=1e30, seed=9, stream_frames=2)
    chunks = list(small_rt.stream(_tokens(8, 80), cfg))
    assert sum(c.size for c in chunks) == 25 * 1920 and len(chunks) >= 10
    ref = small_rt.generate(_tokens(8, 80), pkg.RuntimeGenerateConfig(max_steps=25, eos_threshold=1e30, seed=9))
    assert snr_db(ref.pcm, np.concatenate(chunks)) >= 60.0
    for i, p in enumerate(others):
        assert small_rt.wait(p).n_frames == 12 + i


def test_stream_cancel_mid_flight(pkg, small_tensors):
    rt = pkg.NativeCudaRuntime(tensors=small_tensors, max_batch=2, max_seq_tokens=256, max_frames=96,
                               mimi_group_frames=96)
    try:
        p = rt.submit(_tokens(6, 3), pkg.RuntimeGenerateConfig(max_steps=90, eos_threshold=1e30, stream_frames=1))
        ready, done = rt.read(p, 0)
        assert ready >= 1 and not done
        rt.cancel(p)
        with pytest.raises(pkg.NativeCudaError) as ei:
            rt.wait(p)
        assert ei.value.status == pkg.PTTS_ECANCELED
        assert rt.generate_audio(_tokens(4, 4), pkg.RuntimeGenerateConfig(max_steps=2, eos_threshold=1e30)).shape == (3840,)
    finally:
        rt.close()


def _stream_all(rt, p):
    have, pieces = 0, []
    while True:
        ready, done = rt.read(p, have)
        if ready > have:
            pieces.append(p.pcm[have * 1920: ready * 1920].copy())
            have = ready
        if done:
            return have, pieces


@pytest.mark.parametrize("stream_frames", [1, 5])
def test_stateful_stream_decodes_each_frame_once(pkg, small_tensors, stream_frames):
    """Stateful deliveries (per-request K/V ring of the Mimi transformer): a delivery runs the
    decoder over its new frames + ONE context frame (+ one latent for the upsampling taps), not over the
    18-frame receptive field; 90 frames = 1440 rows wrap the 1024-row ring; the stream still equals the
    whole-chunk decode."""
    rt = pkg.NativeCudaRuntime(tensors=small_tensors, max_batch=4, max_seq_tokens=256, max_frames=96,
                               mimi_group_frames=96)
    try:
        toks = _tokens(9, 5)
        n = 90
        kw = dict(max_steps=n, eos_threshold=1e30, noise=pkg.make_noise(n, 0.7, 6))
        whole = rt.generate(toks, pkg.RuntimeGenerateConfig(**kw))
        before = rt.stage_times()["mimi_frames"]
        p = rt.submit(toks, pkg.RuntimeGenerateConfig(stream_frames=stream_frames, **kw))
        have, pieces = _stream_all(rt, p)
        got = rt.wait(p)
        decoded = rt.stage_times()["mimi_frames"] - before
        assert have == got.n_frames == n
        deliveries = -(-n // stream_frames)
        assert decoded <= n + 2 * deliveries, (decoded, n, deliveries)  # the windowed path re-decodes ~18 per delivery
        stream = np.concatenate(pieces)
        assert snr_db(whole.pcm, stream) >= 60.0, snr_db(whole.pcm, stream)
        np.testing.assert_array_equal(stream, got.pcm)
    finally:
        rt.close()


def test_stateful_streams_share_passes_and_recycle_rings(pkg, small_tensors):
    """More concurrent streams than... rings are per request and recycled: 3 waves of 4 streaming requests next to
    whole-chunk ones; every stream equals its whole-chunk decode."""
    rt = pkg.NativeCudaRuntime(tensors=small_tensors, max_batch=8, max_seq_tokens=256, max_frames=64,
                               mimi_group_frames=128)
    try:
        for wave in range(3):
            cfgs = [dict(max_steps=20 + 3 * i, eos_threshold=1e30, seed=10 * wave + i) for i in range(4)]
            toks = [_tokens(6 + i, 100 + 10 * wave + i) for i in range(4)]
            ps = [rt.submit(toks[i], pkg.RuntimeGenerateConfig(stream_frames=1 + i, **cfgs[i])) for i in range(4)]
            plain = [rt.submit(toks[i], pkg.RuntimeGenerateConfig(**cfgs[i])) for i in range(4)]
            outs = [rt.wait(p) for p in ps]
            refs = [rt.wait(p) for p in plain]
            for o, r in zip(outs, refs):
                assert o.n_frames == r.n_frames
                assert snr_db(r.pcm, o.pcm) >= 60.0, snr_db(r.pcm, o.pcm)
    finally:
        rt.close()


def test_windowed_stream_without_rings(pkg, small_tensors, monkeypatch):
    """PTTS_NO_STREAM_STATE=1: the bounded re-decode window of round 2's first streaming path still works."""
    monkeypatch.setenv("PTTS_NO_STREAM_STATE", "1")
    rt = pkg.NativeCudaRuntime(tensors=small_tensors, max_batch=2, max_seq_tokens=256, max_frames=64,
                               mimi_group_frames=64)
    try:
        toks = _tokens(7, 8)
        kw = dict(max_steps=40, eos_threshold=1e30, seed=3)
        whole = rt.generate(toks, pkg.RuntimeGenerateConfig(**kw))
        before = rt.stage_times()["mimi_frames"]
        p = rt.submit(toks, pkg.RuntimeGenerateConfig(stream_frames=1, **kw))
        _stream_all(rt, p)
        got = rt.wait(p)
        assert rt.stage_times()["mimi_frames"] - before > 5 * 40
        assert snr_db(whole.pcm, got.pcm) >= 60.0
    finally:
        rt.close()


def test_more_streams_than_rings(pkg, small_tensors):
    """40 concurrent streaming requests on an engine with 32 Mimi K/V rings: the first 32 decode statefully, the
    others through the bounded re-decode window; a cancelled stream returns its ring; every stream equals the
    whole-chunk decode of the same request."""
    rt = pkg.NativeCudaRuntime(tensors=small_tensors, max_batch=48, max_seq_tokens=256, max_frames=48,
                               mimi_group_frames=256)
    try:
        toks = [_tokens(5 + i % 9, 300 + i) for i in range(40)]
        cfgs = [dict(max_steps=18 + i % 7, eos_threshold=1e30, seed=500 + i) for i in range(40)]
        ps = [rt.submit(toks[i], pkg.RuntimeGenerateConfig(stream_frames=1 + i % 3, **cfgs[i])) for i in range(40)]
        victim = rt.submit(_tokens(6, 999), pkg.RuntimeGenerateConfig(max_steps=40, eos_threshold=1e30, stream_frames=1))
        rt.read(victim, 0)
        rt.cancel(victim)
        outs = [rt.wait(p) for p in ps]
        with pytest.raises(pkg.NativeCudaError):
            rt.wait(victim)
        refs = [rt.generate(toks[i], pkg.RuntimeGenerateConfig(**cfgs[i])) for i in range(40)]
        for o, r in zip(outs, refs):
            assert o.n_frames == r.n_frames
            assert snr_db(r.pcm, o.pcm) >= 60.0, snr_db(r.pcm, o.pcm)
        # the rings all came back: 32 more streams decode statefully again (3 latent frames per 1-frame delivery at most)
        before = rt.stage_times()["mimi_frames"]
        ps = [rt.submit(toks[i], pkg.RuntimeGenerateConfig(stream_frames=1, **cfgs[i])) for i in range(32)]
        n = sum(rt.wait(p).n_frames for p in ps)
        assert rt.stage_times()["mimi_frames"] - before <= 3 * n
    finally:
        rt.close()
