"""BASELINE.json's configurations at the FULL architecture (6 x 1024 FlowLM, flow depth 6, Mimi
2 x 512 + SEANet) against the CPU restatement (oracle/) on the same seeded inputs:

  C1  "Hello world": 8 x 260 + 3 tokens, forced-EOS schedule, 52 free-running frames
  C2  one 50-token chunk, >= 150 frames: teacher-forced bound on every frame + free-running report
  C3  one batched AR step at B = 64 and B = 300 (ragged prompts), 8 sampled rows against the oracle
  C4  Tv = 125 sin(0.01 i) voice prefix: the donor request and a request on the shared prefix

Tolerances are the north star's: latents |a-b| <= 2e-4 + 1e-3 |b| (fp32 accumulate), waveform
SNR >= 40 dB.  The oracle runs on all host cores (OpenMP) and needs ~4 s for C1, ~15 s for C2.
"""
import json
import os

import numpy as np
import pytest

from conftest import allclose_ref, snr_db

pytestmark = pytest.mark.gpu

ABS, REL = 2e-4, 1e-3
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _tokens(n, seed):
    return np.random.default_rng(seed).integers(0, 4000, n).astype(np.int64)


@pytest.fixture(scope="module")
def full_tensors(ck):
    return ck.make_checkpoint(ck.ArchConfig.full())


@pytest.fixture(scope="module")
def full_oracle(orc, full_tensors):
    orc.set_workers(os.cpu_count() or 1)
    om = orc.OracleModel(full_tensors)
    yield om
    orc.set_workers(1)


@pytest.fixture(scope="module")
def full_rt(pkg, full_tensors):
    rt = pkg.NativeCudaRuntime(tensors=full_tensors, max_batch=320, max_seq_tokens=384, max_frames=192,
                               mimi_group_frames=512)
    yield rt
    rt.close()


def _report(name, **kv):
    """Leaves the measured parity figures where the profile notes pick them up."""
    out = os.path.join(ROOT, "gpurun_out")
    try:
        os.makedirs(out, exist_ok=True)
        with open(os.path.join(out, "parity_fullsize.jsonl"), "a") as f:
            f.write(json.dumps(dict(case=name, **kv)) + "\n")
    except OSError:
        pass


def test_c1_hello_world_full_size(full_rt, full_oracle, pkg):
    wl = __import__("importlib").import_module("pockettts_b200.workload")
    c = wl.hello_world()
    assert c.tokens.tolist() == [260] * 8 + [2994, 578, 263] and c.max_steps == 71
    noise = c.noise()
    kw = dict(max_steps=c.max_steps, eos_threshold=1e30, force_eos_step=c.force_eos_step,
              frames_after_eos=c.frames_after_eos)
    got = full_rt.generate(c.tokens, pkg.RuntimeGenerateConfig(noise=noise, **kw), want_latents=True)
    pcm, lat, eos = full_oracle.generate_audio(c.tokens, noise=noise, **kw)
    assert got.n_frames == lat.shape[0] == 52  # EOS forced at step 46 (= round(11 * 12.5 / 3)), + 5 frames
    err = np.abs(got.latents - lat)
    snr = snr_db(pcm, got.pcm)
    _report("C1", frames=int(got.n_frames), latent_max_abs_err=float(err.max()), pcm_snr_db=float(snr))
    assert allclose_ref(got.latents, lat, ABS, REL), float(err.max())
    assert allclose_ref(got.eos_logits, eos, 5e-4, REL)
    assert snr >= 40.0, snr


def test_c2_long_chunk_teacher_forced_and_free_running(full_rt, full_oracle, pkg):
    """One 50-token chunk run to 160 frames.  Teacher-forced (every step fed the oracle's previous
    latent): every frame inside the latent tolerance.  Free-running (the product path):
    the first frame outside the tolerance and the waveform SNR are reported and bounded."""
    toks = _tokens(50, 501)
    n = 160
    noise = pkg.make_noise(n, 0.7, 502)
    # --- teacher-forced through the stage API, oracle state stepped alongside ---
    s = full_rt.seq_open(toks, reserve_frames=n + 2)
    st = full_oracle.new_state()
    st.prompt(full_oracle.text_embeddings(toks))
    frame = np.full(32, np.nan, np.float32)
    ref = np.empty((n, 32), np.float32)
    worst = 0.0
    try:
        for i in range(n):
            want, want_eos, _ = st.step(frame, noise[i])
            lat, eos = full_rt.step([s], noise[i], frames_in=frame)
            e = np.abs(lat[0] - want)
            worst = max(worst, float((e / (ABS + REL * np.abs(want))).max()))
            assert allclose_ref(lat[0], want, ABS, REL), (i, float(e.max()))
            assert abs(eos[0] - want_eos) <= 5e-4 + REL * abs(want_eos), (i, eos[0], want_eos)
            ref[i] = want
            frame = want
        pcm_tf = s.decode(0, n)
    finally:
        s.close()
    pcm_ref = full_oracle.decode_latents(ref)
    snr_tf = snr_db(pcm_ref, pcm_tf)
    assert snr_tf >= 40.0, snr_tf
    # --- free-running: the scheduler path feeds its own latents back for 160 steps ---
    kw = dict(max_steps=n, eos_threshold=1e30, force_eos_step=-1, frames_after_eos=0)
    got = full_rt.generate(toks, pkg.RuntimeGenerateConfig(noise=noise, **kw), want_latents=True)
    pcm, lat, _ = full_oracle.generate_audio(toks, noise=noise, **kw)
    assert got.n_frames == lat.shape[0] == n
    bad = np.any(np.abs(got.latents - lat) > ABS + REL * np.abs(lat), axis=1)
    first_bad = int(np.argmax(bad)) if bad.any() else -1
    snr = snr_db(pcm, got.pcm)
    _report("C2_chunk160", teacher_forced_worst_frac_of_tol=worst, teacher_forced_pcm_snr_db=float(snr_tf),
            free_running_first_frame_outside_tol=first_bad, free_running_pcm_snr_db=float(snr),
            free_running_latent_max_abs_err=float(np.abs(got.latents - lat).max()))
    assert snr >= 40.0, (snr, first_bad)
    assert first_bad < 0, (first_bad, float(np.abs(got.latents - lat).max()))


@pytest.mark.parametrize("stream_frames", [1, 4])
def test_c2_streamed_per_frame_full_size(full_rt, full_oracle, pkg, stream_frames):
    """C2 as the reference's stream handler delivers it: PCM out while the AR loop runs, the Mimi
    transformer's K/V carried in the request's ring (window 250 rows = 15.6 frames; 160 frames = 2560 rows wrap the
    1024-row ring).  The stream equals the oracle's whole-chunk waveform and the engine's own."""
    toks = _tokens(50, 501)
    n = 160
    noise = pkg.make_noise(n, 0.7, 502)
    kw = dict(max_steps=n, eos_threshold=1e30, force_eos_step=-1, frames_after_eos=0)
    before = full_rt.stage_times()["mimi_frames"]
    p = full_rt.submit(toks, pkg.RuntimeGenerateConfig(noise=noise, stream_frames=stream_frames, **kw))
    have, pieces = 0, []
    while True:
        ready, done = full_rt.read(p, have)
        if ready > have:
            pieces.append(p.pcm[have * 1920: ready * 1920].copy())
            have = ready
        if done:
            break
    got = full_rt.wait(p, want_latents=True)
    decoded = full_rt.stage_times()["mimi_frames"] - before
    assert have == got.n_frames == n and len(pieces) >= n // stream_frames
    assert decoded <= n + 2 * len(pieces) + 2, decoded  # not n * 33
    stream = np.concatenate(pieces)
    whole = full_rt.generate(toks, pkg.RuntimeGenerateConfig(noise=noise, **kw))
    pcm, lat, _ = full_oracle.generate_audio(toks, noise=noise, **kw)
    snr_self, snr = snr_db(whole.pcm, stream), snr_db(pcm, stream)
    _report("C2_streamed", stream_frames=stream_frames, deliveries=len(pieces), mimi_frames_decoded=int(decoded),
            snr_vs_whole_chunk_db=float(snr_self), snr_vs_oracle_db=float(snr))
    assert allclose_ref(got.latents, lat, ABS, REL)
    assert snr_self >= 60.0 and snr >= 40.0, (snr_self, snr)


@pytest.mark.parametrize("batch", [64, 300])
def test_c3_batched_step_full_size(full_rt, full_oracle, pkg, batch):
    """One batched AR step at real shapes: ragged prompts, two teacher-forced steps, the oracle
    checked on 8 sampled rows (the tcgen05 128x32/64/128 tiles and the split-K paths)."""
    rng = np.random.default_rng(batch)
    lens = rng.integers(5, 51, batch)
    sample = sorted(rng.choice(batch, 8, replace=False).tolist())
    seqs, states = [], {}
    try:
        for b in range(batch):
            toks = _tokens(int(lens[b]), 7000 + b)
            seqs.append(full_rt.seq_open(toks, reserve_frames=4))
            if b in sample:
                st = full_oracle.new_state()
                st.prompt(full_oracle.text_embeddings(toks))
                states[b] = st
        frames = np.full((batch, 32), np.nan, np.float32)
        for step in range(2):
            noise = pkg.make_noise(batch, 0.7, 60 + step)
            lat, eos = full_rt.step(seqs, noise, frames_in=frames)
            assert np.isfinite(lat).all()
            nxt = lat.copy()
            for b in sample:
                wl_, we, _ = states[b].step(frames[b], noise[b])
                assert allclose_ref(lat[b], wl_, ABS, REL), (step, b, float(np.abs(lat[b] - wl_).max()))
                assert abs(eos[b] - we) <= 5e-4 + REL * abs(we), (step, b)
                nxt[b] = wl_  # teacher forcing: sampled rows continue from the oracle's latent
            frames = nxt
    finally:
        for s in seqs:
            s.close()


def test_c4_voice_prefix_full_size(full_rt, full_oracle, pkg):
    """Tv = 125 frames of sin(0.01 i) (native_safetensors_integration_test.go:81-84): the first request
    prefills the voice and donates its 7 full pages, the second maps them and prefills 13 voice rows +
    its text.  Both against the oracle."""
    tv = 125
    voice = np.sin(0.01 * np.arange(tv * 1024, dtype=np.float64)).astype(np.float32).reshape(1, tv, 1024)
    ve = pkg.VoiceEmbedding(voice)
    before = full_rt.voice_cache_stats()
    for k, (ntok, seed) in enumerate(((14, 801), (23, 802))):
        toks = _tokens(ntok, seed)
        noise = pkg.make_noise(6, 0.7, 810 + k)
        kw = dict(max_steps=6, eos_threshold=1e30, force_eos_step=3, frames_after_eos=2)
        got = full_rt.generate(toks, pkg.RuntimeGenerateConfig(noise=noise, voice_embedding=ve, **kw), want_latents=True)
        pcm, lat, _ = full_oracle.generate_audio(toks, noise=noise, voice_embedding=ve.rows(), **kw)
        assert got.n_frames == lat.shape[0] == 6
        err = np.abs(got.latents - lat)
        snr = snr_db(pcm, got.pcm)
        _report("C4_donor" if k == 0 else "C4_shared", latent_max_abs_err=float(err.max()), pcm_snr_db=float(snr))
        assert allclose_ref(got.latents, lat, ABS, REL), (k, float(err.max()))
        assert snr >= 40.0, (k, snr)
    after = full_rt.voice_cache_stats()
    assert after["hits"] - before["hits"] == 1 and after["shared_tokens"] - before["shared_tokens"] == 112


def test_f16_and_bf16_load_tensor_widening(pkg, ck, small_tensors):
    """ptts_load_tensor with F16 / BF16 payloads widens exactly as store.go:339-395: an engine fed
    half-precision payloads equals an engine fed the same values already widened to fp32 (subnormal
    halves included)."""
    def run(tensors, dtype):
        rt = pkg.NativeCudaRuntime(max_batch=2, max_seq_tokens=128, max_frames=8, mimi_group_frames=16)
        try:
            for k, a in tensors.items():
                rt.load_tensor(k, a, dtype=dtype)
            rt.finalize()
            cfg = pkg.RuntimeGenerateConfig(max_steps=3, eos_threshold=1e30, seed=5)
            return rt.generate(_tokens(6, 3), cfg, want_latents=True)
        finally:
            rt.close()

    for dtype, npdt in (("f16", np.float16), ("bf16", None)):
        if npdt is not None:
            exact = {k: a.astype(npdt).astype(np.float32) for k, a in small_tensors.items()}
            # subnormal halves (|v| < 6.1e-5) must survive the widening too
            b = exact["flow_lm.flow_net.cond_embed.bias"].copy()
            b[:4] = np.array([5.96e-8, -1.2e-6, 3.0e-5, -6.0e-5], np.float32).astype(np.float16).astype(np.float32)
            exact["flow_lm.flow_net.cond_embed.bias"] = b
        else:
            exact = ck.to_bf16_exact(small_tensors)
        a = run(exact, "f32")
        h = run(exact, dtype)
        np.testing.assert_array_equal(a.latents, h.latents)
        np.testing.assert_array_equal(a.pcm, h.pcm)


def test_bf16_exact_checkpoint_takes_the_two_product_path(pkg, ck, orc, full_tensors, monkeypatch):
    """A checkpoint whose weights are exactly bf16 values (stored as BF16 and widened by the loader,
    store.go:339-395): the weights' lo planes are zero, so the TMA-fed GEMMs skip them — two products per
    k-block, half the weight stream.  Bit-identical to the three-product kernels on the same
    weights (Ahi * 0 adds exact zeros), inside the tolerance against the oracle on the same weights, and the
    kernel classes of a profiled step show the `_bf16w` instantiations."""
    tensors = ck.to_bf16_exact(full_tensors)
    batch = 64
    rng = np.random.default_rng(batch)
    lens = rng.integers(5, 51, batch)
    toks = [_tokens(int(lens[b]), 9000 + b) for b in range(batch)]
    noise = [pkg.make_noise(batch, 0.7, 70 + step) for step in range(2)]

    def run(profile):
        rt = pkg.NativeCudaRuntime(tensors=tensors, max_batch=128, max_seq_tokens=128, max_frames=16,
                                   mimi_group_frames=64)
        seqs = []
        try:
            seqs = [rt.seq_open(t, reserve_frames=4) for t in toks]
            frames = np.full((batch, 32), np.nan, np.float32)
            outs = []
            for step in range(2):
                if profile and step == 1:
                    rt.profile(True)
                lat, eos = rt.step(seqs, noise[step], frames_in=frames)
                outs.append((lat.copy(), eos.copy()))
                frames = lat
            rep = rt.profile_report() if profile else {}
            pcm = seqs[3].decode(0, 2)
            return outs, rep, pcm
        finally:
            for s in seqs:
                s.close()
            rt.close()

    fast, rep, pcm_fast = run(True)
    assert any(k.endswith("_bf16w") for k in rep), sorted(rep)
    assert not any(k.startswith("tcgen05_128x") and not k.endswith("_bf16w") and "ew8" not in k
                   and rep[k].get("launches", 0) > 0 and k in ("tcgen05_128x32", "tcgen05_128x64", "tcgen05_128x128")
                   for k in rep), sorted(rep)
    monkeypatch.setenv("PTTS_NO_BF16_FAST", "1")
    slow, _, pcm_slow = run(False)
    for (la, ea), (lb, eb) in zip(fast, slow):
        np.testing.assert_array_equal(la, lb)
        np.testing.assert_array_equal(ea, eb)
    np.testing.assert_array_equal(pcm_fast, pcm_slow)
    # the oracle on the same (rounded) weights, 4 sampled rows, teacher-forced
    orc.set_workers(os.cpu_count() or 1)
    try:
        om = orc.OracleModel(tensors)
        frames = np.full((batch, 32), np.nan, np.float32)
        states = {}
        for b in (0, 17, 40, 63):
            st = om.new_state()
            st.prompt(om.text_embeddings(toks[b]))
            states[b] = st
        for step in range(2):
            lat, eos = fast[step]
            for b, st in states.items():
                wl_, we, _ = st.step(frames[b], noise[step][b])
                assert allclose_ref(lat[b], wl_, ABS, REL), (step, b, float(np.abs(lat[b] - wl_).max()))
                assert abs(eos[b] - we) <= 5e-4 + REL * abs(we), (step, b)
            frames = lat  # free-running rows: the oracle follows the engine's latents
    finally:
        orc.set_workers(1)
