"""In-process multi-GPU pool: one engine per device behind the same submit / wait calls,
least-loaded dispatch, no collective.  Two engines on ONE device exercise the whole host path on a
single-GPU box; the two-device test runs when the box has them."""
import numpy as np
import pytest

from conftest import allclose_ref, snr_db

pytestmark = pytest.mark.gpu


def _tokens(n, seed):
    return np.random.default_rng(seed).integers(0, 4000, size=n).astype(np.int64)


def _requests(pkg, n):
    out = []
    for i in range(n):
        ms = 4 + i % 6
        out.append((_tokens(3 + (i * 7) % 23, 900 + i),
                    dict(max_steps=ms, eos_threshold=1e30, force_eos_step=ms - 3, frames_after_eos=1,
                         noise=pkg.make_noise(ms, 0.7, 950 + i))))
    return out


def _run_pool(pkg, small_tensors, small_oracle, devices):
    pool = pkg.NativeCudaPool(tensors=small_tensors, devices=devices, max_batch=8, max_seq_tokens=256, max_frames=32,
                              mimi_group_frames=64)
    try:
        assert pool.size == len(devices)
        reqs = _requests(pkg, 24)
        pend = [pool.submit(t, pkg.RuntimeGenerateConfig(**kw)) for t, kw in reqs]
        load = pool.load()
        assert sum(load["dispatched"]) == 24 and min(load["dispatched"]) >= 6, load  # spread over the engines
        outs = [pool.wait(p, want_latents=True) for p in pend]
        for (t, kw), got in zip(reqs, outs):
            pcm, lat, _ = small_oracle.generate_audio(t, **kw)
            assert got.n_frames == lat.shape[0]
            assert allclose_ref(got.latents, lat, 2e-4, 1e-3)
            assert snr_db(pcm, got.pcm) >= 40.0
        after = pool.load()
        assert after["inflight_cost"] == [0] * len(devices) and after["inflight_requests"] == [0] * len(devices)
        # errors keep their own text through the pool
        with pytest.raises(pkg.NativeCudaError, match="out of range"):
            pool.generate_audio([1, 99999], pkg.RuntimeGenerateConfig(max_steps=2))
        p = pool.submit(_tokens(5, 1), pkg.RuntimeGenerateConfig(max_steps=30, eos_threshold=1e30))
        pool.cancel(p)
        with pytest.raises(pkg.NativeCudaError) as ei:
            pool.wait(p)
        assert ei.value.status == pkg.PTTS_ECANCELED
        st = pool.stage_times()
        assert st["steps"] > 0 and st["mimi_frames"] >= sum(o.n_frames for o in outs)
    finally:
        pool.close()
        pool.close()  # idempotent


def test_pool_two_engines_one_device(pkg, small_tensors, small_oracle):
    _run_pool(pkg, small_tensors, small_oracle, [0, 0])


def test_pool_dispatch_prefers_the_idle_engine(pkg, small_tensors):
    """A long request pins engine A; the short ones that follow all go to engine B until the loads cross."""
    pool = pkg.NativeCudaPool(tensors=small_tensors, devices=[0, 0], max_batch=8, max_seq_tokens=512, max_frames=200,
                              mimi_group_frames=256)
    try:
        long = pool.submit(_tokens(40, 5), pkg.RuntimeGenerateConfig(max_steps=180, eos_threshold=1e30))
        first = pool.load()["dispatched"]
        short = [pool.submit(_tokens(4, 10 + i), pkg.RuntimeGenerateConfig(max_steps=3, eos_threshold=1e30)) for i in range(6)]
        d = pool.load()["dispatched"]
        busy = first.index(1)
        assert d[busy] == 1 and d[1 - busy] == 6, d
        for p in short + [long]:
            pool.wait(p)
    finally:
        pool.close()


def test_pool_two_devices(pkg, small_tensors, small_oracle):
    if pkg.lib().ptts_device_count() < 2:
        pytest.skip("needs two GPUs")
    _run_pool(pkg, small_tensors, small_oracle, [0, 1])
