"""Kernel-level parity of the GEMM paths against a float64 numpy reference:
CUDA-core SGEMM / split-K / GEMV (path 1) and the tcgen05 split-bf16 kernel
(path 2).  fp32 kernels: tolerance 1e-5 relative to the row/col norms; the
split-bf16 tensor-core kernel carries ~2^-16 relative error per product, so
its bound is 4e-5 (still 25x inside the 1e-3 latent tolerance)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ACTS = {0: lambda v: v,
        1: lambda v: 0.5 * v * (1 + np.vectorize(__import__("math").erf)(v / np.sqrt(2))),
        2: lambda v: v / (1 + np.exp(-v)),
        3: lambda v: np.where(v > 0, v, np.exp(np.minimum(v, 0)) - 1)}


def _case(M, K, N, seed):
    rng = np.random.default_rng(seed)
    A = rng.standard_normal((M, K)).astype(np.float32)
    W = (rng.standard_normal((N, K)) / np.sqrt(K)).astype(np.float32)
    b = rng.standard_normal(N).astype(np.float32) * 0.1
    return A, W, b


def _check(got, A, W, b, act, tol):
    want = ACTS[act](A.astype(np.float64) @ W.astype(np.float64).T + b)
    scale = np.linalg.norm(A, axis=1, keepdims=True) * np.linalg.norm(W, axis=1)[None, :] + 1.0
    err = np.abs(got - want) / scale
    assert np.isfinite(got).all()
    assert err.max() <= tol, (err.max(), np.unravel_index(err.argmax(), err.shape))


@pytest.mark.parametrize("M,K,N", [(1, 1024, 3072), (3, 32, 1024), (8, 4096, 1024), (20, 1024, 1024),
                                   (65, 1024, 4096), (200, 512, 10240), (256, 4096, 1024), (700, 512, 1536),
                                   (1000, 192, 32), (4097, 768, 128), (333, 3584, 512)])
def test_cuda_core_paths(small_rt, M, K, N):
    A, W, b = _case(M, K, N, M + N)
    _check(small_rt.debug_gemm(A, W, b, act=0, path=1), A, W, b, 0, 1e-5)


@pytest.mark.parametrize("M,K,N", [(128, 64, 128), (128, 1024, 128), (64, 1024, 3072), (200, 512, 10240),
                                   (256, 4096, 1024), (1000, 192, 32), (4097, 768, 128), (333, 3584, 512),
                                   (129, 32, 64), (2000, 128, 256), (777, 1024, 640), (300, 256, 2048),
                                   # > 148 tiles: the persistent loop, TMEM double buffering, ring wrap-around
                                   (40000, 64, 64), (30000, 192, 32), (25000, 256, 256), (21000, 128, 512),
                                   (512, 1024, 3072), (48, 4096, 1024), (16, 512, 512)])
def test_tcgen05_split_bf16_path(small_rt, M, K, N):
    A, W, b = _case(M, K, N, M + N + 1)
    _check(small_rt.debug_gemm(A, W, b, act=0, path=2), A, W, b, 0, 4e-5)


@pytest.mark.parametrize("act", [1, 2, 3])
def test_tcgen05_epilogue_activations(small_rt, act):
    A, W, b = _case(300, 512, 256, act)
    _check(small_rt.debug_gemm(A, W, b, act=act, path=2), A, W, b, act, 4e-5)


def test_tcgen05_is_deterministic(small_rt):
    A, W, b = _case(500, 1024, 512, 77)
    x = small_rt.debug_gemm(A, W, b, path=2)
    y = small_rt.debug_gemm(A, W, b, path=2)
    np.testing.assert_array_equal(x, y)


def test_split_k_is_deterministic(small_rt):
    A, W, b = _case(100, 4096, 1024, 78)
    x = small_rt.debug_gemm(A, W, b, path=1)
    y = small_rt.debug_gemm(A, W, b, path=1)
    np.testing.assert_array_equal(x, y)


# ---- TMA-fed A operand (bf16 hi/lo planes written by the producer, both operands by TMA) ----
@pytest.mark.parametrize("M,K,N", [(16, 32, 1024), (300, 1024, 3072), (464, 4096, 1024), (25000, 512, 1536),
                                   (40000, 32, 64), (9, 512, 32), (2000, 2048, 512)])
def test_tma_fed_dense(small_rt, M, K, N):
    A, W, b = _case(M, K, N, M + N + 2)
    got, sp = small_rt.debug_gemm_tma(A, W, b, want_split=True)
    _check(got, A, W, b, 0, 4e-5)
    # the split copy carries ~16 mantissa bits of the fp32 output
    assert np.abs(sp - got).max() <= 2.0 ** -15 * max(1.0, np.abs(got).max())


@pytest.mark.parametrize("M,cin,taps,N", [(5000, 64, 3, 32), (3000, 128, 3, 64), (4097, 128, 2, 256),
                                          (700, 512, 7, 512), (1500, 256, 2, 640), (20000, 32, 1, 64)])
def test_tma_fed_causal_conv_form(small_rt, M, cin, taps, N):
    """Implicit-GEMM conv over overlapping rows through the 3-D TMA map {cin, taps, rows}:
    out[m] = sum_t W[:, t*cin:(t+1)*cin] @ x[m - (taps-1) + t], rows before 0 are zero."""
    rng = np.random.default_rng(M + N)
    x = rng.standard_normal((M, cin)).astype(np.float32)
    W = (rng.standard_normal((N, taps * cin)) / np.sqrt(taps * cin)).astype(np.float32)
    b = rng.standard_normal(N).astype(np.float32) * 0.1
    xp = np.concatenate([np.zeros((taps - 1, cin), np.float32), x])
    im2col = np.concatenate([xp[t:t + M] for t in range(taps)], axis=1)  # [M, taps*cin]
    got = small_rt.debug_gemm_tma(x, W, b, act=3, taps=taps)
    _check(got, im2col, W, b, 3, 4e-5)


@pytest.mark.parametrize("M,K,N,bn,S", [(256, 1024, 3072, 256, 4), (256, 1024, 1024, 256, 8), (200, 4096, 1024, 128, 8),
                                        (64, 1024, 4096, 256, 3), (500, 512, 512, 128, 6), (1000, 1024, 1024, 256, 2),
                                        (300, 512, 10240, 256, 1), (90, 1024, 512, 64, 5), (256, 4096, 1024, 32, 7)])
def test_cluster_split_k_distributed_reduction(small_rt, M, K, N, bn, S):
    """Split-K through a thread-block cluster: the S CTAs of a tile exchange row bands of their partial
    accumulators through DSMEM and each reduces + stores its own band (any S <= 8, any tile width);
    results match the fp64 reference, are deterministic, and equal the epilogue's split-plane copy."""
    A, W, b = _case(M, K, N, M + N + S)
    small_rt.tc_force(bn, S)
    try:
        got, sp = small_rt.debug_gemm_tma(A, W, b, act=1, want_split=True)
        again = small_rt.debug_gemm_tma(A, W, b, act=1)
    finally:
        small_rt.tc_force(0, 0)
    _check(got, A, W, b, 1, 4e-5)
    np.testing.assert_array_equal(got, again)
    assert np.abs(sp - got).max() <= 2.0 ** -15 * max(1.0, np.abs(got).max())
