"""Model-level behaviour of the CPU restatement on a tiny random checkpoint
(CPU only; these pin the oracle's driver semantics to the reference's rules)."""
import numpy as np
import pytest

from conftest import allclose_ref


def _noise(pkg, n, seed):
    return pkg.make_noise(n, 0.7, seed)


def test_eos_countdown_rule(tiny_model, pkg):
    # runtime_native_safetensors.go:178-190: frames = eos_step + 1 + frames_after_eos, capped
    n = _noise(pkg, 30, 1)
    for eos_step, fae, cap, want in ((4, 3, 30, 8), (0, 5, 30, 6), (10, 3, 12, 12), (-1, 3, 9, 9), (2, 0, 30, 3)):
        _, lat, _ = tiny_model.generate_audio([1, 2, 3], noise=n, max_steps=cap, eos_threshold=1e30,
                                              force_eos_step=eos_step, frames_after_eos=fae, want_pcm=False)
        assert lat.shape[0] == want, (eos_step, fae, cap)


def test_natural_eos_uses_raw_logit_threshold(tiny_model, pkg):
    n = _noise(pkg, 12, 2)
    _, _, eos = tiny_model.generate_audio([4, 5], noise=n, max_steps=12, eos_threshold=1e30, want_pcm=False)
    k = int(np.argmax(eos[:6]))
    thr = float(np.sort(eos[:6])[-2] + eos[k]) / 2
    _, lat, _ = tiny_model.generate_audio([4, 5], noise=n, max_steps=12, eos_threshold=thr, frames_after_eos=1,
                                          want_pcm=False)
    assert lat.shape[0] == k + 2


def test_guards(tiny_model, orc, pkg):
    n = _noise(pkg, 4, 3)
    with pytest.raises(orc.OracleError, match="must not be empty"):
        tiny_model.generate_audio([], noise=n, max_steps=2)
    with pytest.raises(orc.OracleError, match="out of range"):
        tiny_model.generate_audio([99999], noise=n, max_steps=2)
    st = tiny_model.new_state()
    with pytest.raises(orc.OracleError, match="mutually exclusive"):
        tiny_model.generate_audio([1], noise=n, max_steps=2, state=st, voice_embedding=np.zeros((1, tiny_model.D)))


def test_stateful_steps_equal_one_shot_prefill(tiny_model):
    """KV append/growth (flow_transformer.go:32-108): prompting token by token gives
    the same cache as one prefill (flow_guards_test.go:207-257, :366-461)."""
    emb = tiny_model.text_embeddings([3, 1, 4, 1, 5, 9, 2])
    a = tiny_model.new_state()
    a.prompt(emb)
    b = tiny_model.new_state()
    for i in range(emb.shape[0]):
        b.prompt(emb[i:i + 1])
    assert a.offset() == b.offset() == 7
    assert b.capacity() == 8  # doubling: 1,2,4,8
    for layer in range(tiny_model.L):
        for which in (0, 1):
            assert allclose_ref(b.read_kv(layer, which), a.read_kv(layer, which), 1e-5, 1e-5)


def test_projector_equals_denorm_plus_conv(tiny_model):
    # model_decode_test.go:174-272
    lat = np.random.default_rng(0).standard_normal((5, 32)).astype(np.float32)
    np.testing.assert_allclose(tiny_model.latent_to_mimi(lat), tiny_model.latent_to_mimi(lat, unfused=True),
                               rtol=1e-5, atol=1e-5)


def test_mimi_decode_shapes_and_causality(tiny_model):
    lat = np.random.default_rng(1).standard_normal((6, 32)).astype(np.float32)
    whole = tiny_model.decode_latents(lat)
    assert whole.shape == (6 * 1920,) and np.isfinite(whole).all()
    for n in (1, 3):
        np.testing.assert_allclose(tiny_model.decode_latents(lat[:n]), whole[: n * 1920], rtol=1e-4, atol=1e-5)


def test_voice_state_slots_past_offset_are_ignored(tiny_model, pkg):
    """flow_transformer.go:404-420: cache slots >= offset get key position -1, so
    whatever they hold (NaN in upstream voice files) never reaches the output."""
    rng = np.random.default_rng(2)
    n = _noise(pkg, 3, 4)
    base = [rng.standard_normal((2, 1, 6, tiny_model.H, tiny_model.Dh)).astype(np.float32)
            for _ in range(tiny_model.L)]
    outs = []
    for fill in (np.nan, 123.0):
        st = tiny_model.new_state()
        for layer, c in enumerate(base):
            c = c.copy()
            c[:, :, 4:] = fill
            st.set_voice_layer(layer, c, 4)
        _, lat, _ = tiny_model.generate_audio([7, 8], noise=n, max_steps=3, eos_threshold=1e30, state=st,
                                              want_pcm=False)
        assert np.isfinite(lat).all()
        outs.append(lat)
    np.testing.assert_array_equal(outs[0], outs[1])


def test_lsd_steps_change_result_but_stay_finite(tiny_model):
    c = np.random.default_rng(3).standard_normal(tiny_model.D).astype(np.float32)
    x0 = np.random.default_rng(4).standard_normal(32).astype(np.float32)
    a, b = tiny_model.lsd_decode(c, x0, 1), tiny_model.lsd_decode(c, x0, 4)
    assert np.isfinite(a).all() and np.isfinite(b).all() and np.abs(a - b).max() > 1e-6
    # one Euler step == x0 + v(c, 0, 1, x0)
    np.testing.assert_allclose(a, x0 + tiny_model.flow_direction(c, 0.0, 1.0, x0), rtol=1e-6, atol=1e-6)
