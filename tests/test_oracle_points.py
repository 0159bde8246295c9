"""oracle/points.py (numpy restatement of the reference's Halton point sampler) against the known-answer vectors the
reference's own ext/halton/halton.h produced (tests/golden/make_golden_points.py), bit for bit."""
import os

import numpy as np
import pytest

from oracle.points import Halton, Lrand48

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "halton.npz")


@pytest.mark.parametrize("frame", [0, 1, 2, 1234567])
def test_halton_matches_reference(frame):
    z = np.load(GOLDEN)
    got = Halton(frame).sample(z[f"f{frame}_dim"], z[f"f{frame}_index"])
    assert np.array_equal(got.view("u4"), z[f"f{frame}_value"].view("u4"))
    assert got.min() >= 0.0 and got.max() < 1.0


def test_lrand48_known_answers():
    """glibc: srand48(0); lrand48() x3 -> 366850414, 1610402240, 206956554"""
    r = Lrand48(0)
    assert [r(), r(), r()] == [366850414, 1610402240, 206956554]


def test_halton_matches_live_reference():
    """where the reference is compiled (build container): more samples, all 256 dimensions"""
    import ctypes as C
    so = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "libref_halton.so")
    if not os.path.exists(so):
        pytest.skip("oracle/_ref not built here")
    L = C.CDLL(so)
    L.ref_halton_sample.argtypes = [C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64]
    rng = np.random.default_rng(5)
    n = 100000
    idx = rng.integers(0, 2**33, n, dtype=np.uint64)
    dim = rng.integers(0, 256, n).astype(np.int32)
    want = np.zeros(n, np.float32)
    L.ref_halton_sample(3, idx.ctypes.data, dim.ctypes.data, want.ctypes.data, n)
    assert np.array_equal(Halton(3).sample(dim, idx).view("u4"), want.view("u4"))
