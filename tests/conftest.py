import importlib
import os
import sys

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)
GOLDEN = os.path.join(REPO, "tests", "golden")
PKG_NAME = "variational-self-organizing-maps_b200"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu under gpurun)")


@pytest.fixture(scope="session")
def vsom():
    """The product package (ctypes binding of libvsom_b200.so)."""
    return importlib.import_module(PKG_NAME)


@pytest.fixture(scope="session")
def po():
    """The CPU checkers (oracle/pyoracle.py) — test infrastructure only."""
    from oracle import pyoracle

    pyoracle.Oracle.lib()
    return pyoracle


def bits(a):
    a = np.ascontiguousarray(a)
    if a.dtype == np.float32:
        return a.view(np.uint32)
    if a.dtype == np.float64:
        return a.view(np.uint64)
    return a


def assert_bit_equal(a, b, what=""):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape, f"{what}: shape {a.shape} vs {b.shape}"
    assert a.dtype == b.dtype, f"{what}: dtype {a.dtype} vs {b.dtype}"
    ba, bb = bits(a), bits(b)
    if not np.array_equal(ba, bb):
        bad = np.argwhere(ba != bb)
        i = tuple(bad[0])
        raise AssertionError(f"{what}: {len(bad)} of {a.size} entries differ; first at {i}: {a[i]!r} vs {b[i]!r}")


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)
