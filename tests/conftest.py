import json
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def cube():
    """The reference's test/mesh/cube fixture, decoded by tests/golden/make_cube_fixture.py."""
    d = json.load(open(os.path.join(ROOT, "tests", "golden", "cube_fixture.json")))
    s = d["serial"]
    s["coords"] = np.array(s["coords"])
    s["tets"] = np.array(s["tets"], dtype=np.int32)
    s["parts"] = d["parts"]
    return s


@pytest.fixture(scope="session")
def hostcheck():
    import ctypes as C
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "tests", "hostcheck"), "-s"])
    return C.CDLL(os.path.join(ROOT, "tests", "hostcheck", "libhostcheck.so"))


@pytest.fixture(scope="session")
def gxlib():
    so = os.path.join(ROOT, "goal_b200", "libgoal_b200.so")
    if not os.path.exists(so):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "goal_b200", "csrc"), "-s"])
    from goal_b200.binding import load_library
    return load_library()


def relerr(a, b):
    """norm-wise relative error max|a-b| / max|b| (the 1e-12 parity bar of BASELINE.md 5)."""
    a, b = np.asarray(a), np.asarray(b)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


def blockerr(A, B, bs=4):
    """Largest relative error of any bs x bs block: max over blocks of max|A_blk - B_blk| / max|B_blk|.  Stricter than
    relerr: entries orders of magnitude below the global maximum are still constrained by their own block's scale."""
    A, B = np.asarray(A, dtype=np.float64), np.asarray(B, dtype=np.float64)
    n0, n1 = A.shape[0] // bs, A.shape[1] // bs
    a = A[:n0 * bs, :n1 * bs].reshape(n0, bs, n1, bs)
    b = B[:n0 * bs, :n1 * bs].reshape(n0, bs, n1, bs)
    num = np.abs(a - b).max(axis=(1, 3))
    den = np.abs(b).max(axis=(1, 3))
    ok = den > 0
    return float((num[ok] / den[ok]).max()) if ok.any() else 0.0
