"""Pins the CPU restatement (oracle/) to the reference's own known-answer tests.

Every case reproduces a test of /root/reference (file:line cited) with the same
inputs and expected values, so the oracle the GPU is compared against is
anchored to the reference rather than to itself."""
import numpy as np
import pytest

from conftest import seq_data


def test_dot_kats(orc):
    # internal/runtime/tensor/dot_test.go:9-31
    assert orc.dot([3], [4]) == 12.0
    assert orc.dot([1, 2, 3], [4, 5, 6]) == 32.0
    assert orc.dot([-1, 2], [3, -4]) == -11.0
    assert orc.dot([0, 0, 0], [1, 2, 3]) == 0.0
    assert orc.dot([1] * 16, [2] * 16) == 32.0  # "long enough for AVX2"


@pytest.mark.parametrize("n", [1, 7, 8, 9, 31, 32, 33, 40, 63, 64, 100, 1024, 1027])
def test_dot_avx2_matches_float64(orc, n):
    # dot_amd64.s:34-116 vs dot.go:10-38: both orders agree with an fp64 sum to fp32 rounding
    rng = np.random.default_rng(n)
    a = rng.standard_normal(n).astype(np.float32)
    b = rng.standard_normal(n).astype(np.float32)
    want = float(np.dot(a.astype(np.float64), b.astype(np.float64)))
    assert abs(orc.dot(a, b) - want) <= 1e-5 * max(1.0, np.sum(np.abs(a * b)))


def test_axpy_kats(orc):
    # internal/runtime/tensor/axpy_test.go:9-63
    np.testing.assert_allclose(orc.axpy([1, 2, 3], 0.5, [4, 5, 6]), [3, 4.5, 6], atol=1e-5)
    np.testing.assert_allclose(orc.axpy([1, 2, 3, 4], 2.0, [10, 20]), [21, 42, 3, 4], atol=1e-5)
    np.testing.assert_array_equal(orc.axpy([1, 2, 3], 0.0, [9, 9, 9]), [1, 2, 3])


def test_softmax_kat(orc):
    # internal/runtime/tensor/nn_ops_test.go:9-25
    got = orc.softmax([1, 2, 3])
    np.testing.assert_allclose(got, [0.09003057, 0.24472848, 0.66524094], atol=1e-6)


def test_layernorm_kat(orc):
    # internal/runtime/tensor/nn_ops_test.go:27-43
    got = orc.layernorm([1, 2, 3, 4], eps=1e-5)
    np.testing.assert_allclose(got, [-1.3416355, -0.44721183, 0.44721183, 1.3416355], atol=1e-5)


def test_linear_kat(orc):
    # internal/runtime/tensor/nn_ops_test.go:82-100
    got = orc.linear(np.array([[1, 2], [3, 4]]), np.array([[1, 0], [0, 1], [1, 1]]), [0.5, -0.5, 1])
    np.testing.assert_array_equal(got.ravel(), [1.5, 1.5, 4, 3.5, 3.5, 8])


def test_linear_workers_equal(orc):
    # internal/native/linear_into_test.go:11-91 (parallel == sequential)
    x = seq_data(3 * 64).reshape(3, 64)
    w = seq_data(520 * 64).reshape(520, 64)
    b = seq_data(520)
    a = orc.linear(x, w, b)
    orc.set_workers(4)
    try:
        c = orc.linear(x, w, b)
        d = orc.linear(x[:1], np.tile(w, (9, 1)), np.tile(b, 9))
    finally:
        orc.set_workers(1)
    np.testing.assert_array_equal(a, c)
    np.testing.assert_array_equal(d[0, :520], a[0])


def test_rope_kat(orc):
    # internal/runtime/ops/rope_test.go:10-24
    x = np.array([1, 0, 0, 1], np.float32).reshape(1, 2, 2)
    got = orc.rope(x, np.zeros((2, 1)), np.ones((2, 1)), 0)
    np.testing.assert_allclose(got.ravel(), [0, 1, -1, 0], atol=1e-6)
    with pytest.raises(orc.OracleError):
        orc.rope(x, np.zeros((2, 1)), np.ones((2, 1)), 1)


def test_attention_causal_kat(orc):
    # internal/runtime/ops/attention_test.go:33-51
    q = np.array([1, 1], np.float32).reshape(1, 1, 2, 1)
    k = np.array([0, 10], np.float32).reshape(1, 1, 2, 1)
    v = np.array([1, 5], np.float32).reshape(1, 1, 2, 1)
    got = orc.attention_causal(q, k, v, True, 0).ravel()
    assert abs(got[0] - 1.0) <= 1e-4
    assert got[1] > 4.0


def test_attention_positions_context_invalid_keys_kat(orc):
    # internal/runtime/ops/attention_test.go:53-71
    q = np.zeros((1, 1, 2, 1), np.float32)
    k = np.zeros((1, 1, 4, 1), np.float32)
    v = np.array([100, 1, 3, 20], np.float32).reshape(1, 1, 4, 1)
    got = orc.attention_positions(q, k, v, [2, 3], [-1, 1, 2, 3], 2).ravel()
    assert abs(got[0] - 2.0) <= 1e-4
    assert abs(got[1] - 11.5) <= 1e-4


def test_attention_positions_matches_causal_offset(orc):
    # internal/runtime/ops/attention_test.go:73-91
    q = seq_data(1 * 2 * 3 * 4).reshape(1, 2, 3, 4)
    k = seq_data(1 * 2 * 5 * 4).reshape(1, 2, 5, 4)
    v = seq_data(1 * 2 * 5 * 3).reshape(1, 2, 5, 3)
    got = orc.attention_positions(q, k, v, [2, 3, 4], [0, 1, 2, 3, 4], -1)
    want = orc.attention_causal(q, k, v, True, 2)
    np.testing.assert_allclose(got, want, atol=1e-4)


def test_attention_masked_nan_keys_are_skipped(orc):
    # attention.go:403-406: masked slots are skipped, not multiplied (NaN-safe)
    q = np.ones((1, 1, 1, 2), np.float32)
    k = np.array([[np.nan, np.nan], [1, 0]], np.float32).reshape(1, 1, 2, 2)
    v = np.array([[np.nan, np.nan], [3, 4]], np.float32).reshape(1, 1, 2, 2)
    got = orc.attention_positions(q, k, v, [1], [-1, 1], -1).ravel()
    np.testing.assert_allclose(got, [3, 4], atol=1e-6)
    # all-masked row -> zeros (attention.go:417-419)
    got = orc.attention_positions(q, k, v, [0], [-1, 1], -1).ravel()
    np.testing.assert_array_equal(got, [0, 0])


def test_conv1d_kats(orc):
    # internal/runtime/ops/conv1d_test.go:9-23, 51-75
    x = np.array([1, 2, 3, 4], np.float32).reshape(1, 1, 4)
    got = orc.conv1d(x, np.ones((1, 1, 2), np.float32))
    np.testing.assert_array_equal(got.ravel(), [3, 5, 7])
    x2 = np.array([1, 2, 3, 4, 10, 20, 30, 40], np.float32).reshape(1, 2, 4)
    got = orc.conv1d(x2, np.ones((2, 1, 2), np.float32), groups=2)
    np.testing.assert_array_equal(got.ravel(), [3, 5, 7, 30, 50, 70])


def test_conv1d_leftpad_matches_explicit_prepend(orc):
    # internal/runtime/ops/conv1d_test.go:77-118
    x = np.array([1, 2, 3, 4, 10, 20, 30, 40], np.float32).reshape(1, 2, 4)
    w = np.array([1, 1, 1, 1, 1, 1, 2, 2, 2, 2, 2, 2], np.float32).reshape(2, 2, 3)
    b = np.array([0.25, -0.5], np.float32)
    got = orc.conv1d(x, w, b, stride=2, lpad=2)
    padded = np.concatenate([np.zeros((1, 2, 2), np.float32), x], axis=2)
    want = orc.conv1d(padded, w, b, stride=2)
    np.testing.assert_allclose(got, want, atol=1e-5)


def test_conv1d_parallel_equals_sequential(orc):
    # internal/runtime/ops/conv1d_test.go:25-49
    x = seq_data(16 * 64).reshape(1, 16, 64)
    w = seq_data(32 * 16 * 3).reshape(32, 16, 3)
    b = seq_data(32)
    want = orc.conv1d(x, w, b, lpad=1, rpad=1)
    orc.set_workers(4)
    try:
        got = orc.conv1d(x, w, b, lpad=1, rpad=1)
    finally:
        orc.set_workers(1)
    np.testing.assert_allclose(got, want, atol=1e-4)


def test_convtranspose1d_kats(orc):
    # internal/runtime/ops/convtranspose1d_test.go:10-24
    x = np.array([1, 2, 3], np.float32).reshape(1, 1, 3)
    got = orc.convtranspose1d(x, np.ones((1, 1, 2), np.float32))
    np.testing.assert_array_equal(got.ravel(), [1, 3, 5, 3])
    # grouped with bias (:167-193)
    x = np.arange(1, 9, dtype=np.float32).reshape(1, 4, 2)
    w = np.array([1, 10, 100, 1000], np.float32).reshape(4, 1, 1)
    got = orc.convtranspose1d(x, w, [1, 2], groups=2)
    np.testing.assert_array_equal(got.ravel(), [32, 43, 7502, 8602])
    # depthwise (:195-221)
    x = np.array([1, 2, 3, 4, 0, 6], np.float32).reshape(1, 2, 3)
    w = np.array([1, 1, 2, 0], np.float32).reshape(2, 1, 2)
    got = orc.convtranspose1d(x, w, [0.5, -0.5], groups=2)
    np.testing.assert_array_equal(got.ravel(), [1.5, 3.5, 5.5, 3.5, 7.5, -0.5, 11.5, -0.5])


def test_repack_convtranspose_kernel_kat(orc):
    # internal/runtime/ops/convtranspose1d_test.go:50-74
    w = np.arange(1, 13, dtype=np.float32).reshape(2, 3, 2)
    got = orc.repack_convtr(w).ravel()
    np.testing.assert_array_equal(got, [1, 7, 3, 9, 5, 11, 2, 8, 4, 10, 6, 12])


def test_convtranspose1d_right_trim_matches_narrow(orc):
    # internal/runtime/ops/convtranspose1d_test.go:118-165
    for (ci, co, k, L, trim) in ((3, 4, 4, 5, 2), (2, 3, 5, 6, 3)):
        x = seq_data(ci * L).reshape(1, ci, L)
        w = seq_data(ci * co * k).reshape(ci, co, k)
        b = seq_data(co)
        got = orc.convtranspose1d(x, w, b, stride=2, right_trim=trim)
        full = orc.convtranspose1d(x, w, b, stride=2)
        np.testing.assert_allclose(got, full[:, :, : full.shape[2] - trim], atol=1e-5)


def test_mlp_silu_kat(orc):
    # internal/runtime/ops/mlp_test.go:9-28
    got = orc.mlp_silu(np.array([[1, -1]]), np.eye(2), None, np.ones((1, 2)), None)
    assert abs(float(got.ravel()[0]) - 0.4621172) <= 1e-4


def test_activations(orc):
    # internal/native/tensor_util.go:73-128
    x = np.array([-2.0, -0.5, 0.0, 0.5, 3.0], np.float32)
    xd = x.astype(np.float64)
    from math import erf
    np.testing.assert_allclose(orc.silu(x), xd / (1 + np.exp(-xd)), rtol=1e-6)
    np.testing.assert_allclose(orc.gelu(x), [0.5 * v * (1 + erf(v / np.sqrt(2))) for v in xd], rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(orc.elu(x), np.where(xd <= 0, np.exp(xd) - 1, xd), rtol=1e-6, atol=1e-7)


def test_rmsnorm_alpha_uses_unbiased_variance_without_centering(orc):
    # internal/native/tensor_util.go:273-326
    x = np.array([[1.0, 2.0, 3.0, 6.0]], np.float32)
    alpha = np.array([1.0, 2.0, 0.5, 1.0], np.float32)
    var = np.var(x.astype(np.float64), ddof=1)
    want = x * (1.0 / np.sqrt(var + 1e-5)) * alpha
    np.testing.assert_allclose(orc.rmsnorm_alpha(x, alpha), want, rtol=1e-6)


def test_voice_kv_layout_kat(orc):
    # internal/native/flow_guards_test.go:259-332: [2,1,T=2,H=2,D=1]
    cache = np.array([1, 2, 3, 4, 5, 6, 7, 8], np.float32).reshape(2, 1, 2, 2, 1)
    k, v = orc.split_voice_kv(cache)
    np.testing.assert_array_equal(k.ravel(), [1, 3, 2, 4])
    np.testing.assert_array_equal(v.ravel(), [5, 7, 6, 8])


def test_mimi_window_attention_kat(orc):
    # internal/native/model_decode_test.go:99-159: in_proj routes x to V only,
    # q = k = 0, context 2 -> each row averages itself and its predecessor.
    x = np.array([[1, 10], [3, 30], [20, 200]], np.float32)
    v = x.reshape(1, 1, 3, 2)
    z = np.zeros_like(v)
    pos = [0, 1, 2]
    got = orc.attention_positions(z, z, v, pos, pos, 2).reshape(3, 2)
    np.testing.assert_allclose(got, [[1, 10], [2, 20], [11.5, 115]], atol=1e-5)


def test_stateful_attention_matches_full_last_token(orc):
    # internal/native/flow_guards_test.go:366-461 (identity projections, 1 head of dim 2)
    x = np.array([[1, 0], [0, 1], [1, 1]], np.float32)
    cos = np.array([[1], [0], [-1]], np.float32)
    sin = np.array([[0], [1], [0]], np.float32)
    q = orc.rope(x.reshape(1, 3, 2), cos, sin, 0).reshape(1, 1, 3, 2)
    full = orc.attention_causal(q, q, x.reshape(1, 1, 3, 2), True, 0)
    # incremental: cache holds rotated keys of the 2-token prefix, then the step token
    kq_last = orc.rope(x[2:].reshape(1, 1, 2), cos, sin, 2).reshape(1, 1, 1, 2)
    kcache = np.concatenate([q[:, :, :2], kq_last], axis=2)
    got = orc.attention_positions(kq_last, kcache, x.reshape(1, 1, 3, 2), [2], [0, 1, 2], -1)
    np.testing.assert_allclose(got.ravel(), full[0, 0, 2], atol=1e-5)


def test_estimate_max_frames(orc):
    # internal/text/prepare_test.go (EstimateMaxFrames = ceil((n/3+2)*12.5))
    assert orc.estimate_max_frames(0) == 25
    assert orc.estimate_max_frames(3) == 38
    assert orc.estimate_max_frames(11) == 71
    assert orc.estimate_max_frames(50) == 234
    assert orc.estimate_max_frames(-5) == 25
