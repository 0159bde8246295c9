"""Known-answer vectors for pointsampler() in Halton mode, produced by the reference's own ext/halton/halton.h
(oracle/_ref/libref_halton.so, compiled in place by oracle/Makefile).  Build container only:

    python tests/golden/make_golden_points.py   ->  tests/golden/halton.npz
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))

if __name__ == "__main__":
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "ref"])
    L = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libref_halton.so"))
    L.ref_halton_sample.argtypes = [C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64]
    rng = np.random.default_rng(13)
    pack = {}
    edge = np.array([0, 1, 2, 3, 255, 256, 65535, 65536, 2**24 - 1, 2**24, 2**31 - 1, 2**31, 2**32 - 1, 2**32, 2**32 + 5, 2**40 + 77], np.uint64)
    for frame in (0, 1, 2, 1234567):
        n = 6000
        idx = np.concatenate([rng.integers(0, 2**32, n // 3, dtype=np.uint64), rng.integers(0, 200000, n // 3, dtype=np.uint64),
                              rng.integers(2**32, 2**36, n // 3, dtype=np.uint64), np.repeat(edge, 256)])
        dim = np.concatenate([rng.integers(0, 256, n).astype(np.int32), np.tile(np.arange(256, dtype=np.int32), len(edge))])
        out = np.zeros(len(idx), np.float32)
        L.ref_halton_sample(frame, idx.ctypes.data, dim.ctypes.data, out.ctypes.data, len(idx))
        pack[f"f{frame}_index"], pack[f"f{frame}_dim"], pack[f"f{frame}_value"] = idx, dim, out
    pack["frames"] = np.int64([0, 1, 2, 1234567])
    path = os.path.join(HERE, "halton.npz")
    np.savez_compressed(path, **pack)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")
