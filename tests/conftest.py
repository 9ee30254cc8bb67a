import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def golden(name):
    return np.load(os.path.join(GOLDEN, name))


def bits(a):
    """float32 array -> uint32 view, for bit-exact comparisons that treat NaN payloads as data."""
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def assert_same_bits(a, b, what=""):
    a = np.ascontiguousarray(a)
    b = np.ascontiguousarray(b)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    if a.tobytes() != b.tobytes():
        av = a.view(np.uint8).reshape(-1)
        bv = b.view(np.uint8).reshape(-1)
        bad = np.nonzero(av != bv)[0]
        raise AssertionError(f"{what}: {bad.size} bytes differ, first at byte {bad[0]}")


@pytest.fixture(scope="session")
def q_default():
    return golden("q_golden.npz")["q"][0]
