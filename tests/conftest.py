import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import __graft_entry__ as ge  # noqa: E402


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu under gpurun)")


@pytest.fixture(scope="session")
def pkg():
    return ge.package()


@pytest.fixture(scope="session")
def ck():
    ge.package()
    import importlib
    return importlib.import_module("pockettts_b200.checkpoint")


@pytest.fixture(scope="session")
def orc():
    """The CPU restatement (test infrastructure only)."""
    o = ge.oracle()
    o.build()
    return o


@pytest.fixture(scope="session")
def tiny_tensors(ck):
    return ck.make_checkpoint(ck.ArchConfig.tiny(), seed=11)


@pytest.fixture(scope="session")
def tiny_model(orc, tiny_tensors):
    return orc.OracleModel(tiny_tensors)


@pytest.fixture(scope="session")
def small_tensors(ck):
    return ck.make_checkpoint(ck.ArchConfig.small(), seed=7)


@pytest.fixture(scope="session")
def small_oracle(orc, small_tensors):
    return orc.OracleModel(small_tensors)


@pytest.fixture(scope="session")
def small_rt(pkg, small_tensors):
    rt = pkg.NativeCudaRuntime(tensors=small_tensors, max_batch=16, max_seq_tokens=512, max_frames=96,
                               mimi_group_frames=128)
    yield rt
    rt.close()


def seq_data(n):
    """ops/test_helpers_test.go:11-18 (seqDataT)."""
    i = np.arange(n)
    return (((i % 17) - 8) / 17).astype(np.float32)


def allclose_ref(got, want, abs_tol=2e-4, rel_tol=1e-3):
    """|a-b| <= abs + rel*|b| per element (SURVEY §7 'hard parts')."""
    got = np.asarray(got, dtype=np.float64)
    want = np.asarray(want, dtype=np.float64)
    return bool(np.all(np.abs(got - want) <= abs_tol + rel_tol * np.abs(want)))


def snr_db(ref, got):
    ref = np.asarray(ref, dtype=np.float64)
    got = np.asarray(got, dtype=np.float64)
    noise = float(np.sum((ref - got) ** 2))
    sig = float(np.sum(ref ** 2))
    return 10.0 * np.log10(sig / max(noise, 1e-300))
