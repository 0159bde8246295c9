"""Known-answer vectors for accel_closest (include/accel.h:47, src/accel.d/qbvhmp.c:1493-1600) from the unmodified reference
(oracle/_ref/libcorona_ref.so), on the scenes + reference-built trees of the existing traversal fixtures.  Build container only:

    python tests/golden/make_golden_closest.py   ->  tests/golden/closest.npz
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import Golden, GOLDEN_NAMES, R   # noqa: E402
from oracle.binding import Ref, build         # noqa: E402


def queries(g, n, seed):
    """rays of the fixture; centre somewhere along each ray (a seventh of them exactly on the first surface: the
    "straight hit" exit), search limit 2*centre (the two-sided interval halfvec.h sets up), min_dist 0"""
    rng = np.random.default_rng(seed)
    rays = np.concatenate([g.rays, g.bounce])[:n].copy()
    hits = np.concatenate([g.hits, g.hits_bounce])[:n]
    d = hits["dist"].copy()
    d[R.hit_prim64(hits) == R.INVALID_PRIMID] = 5.0
    centre = (d * rng.uniform(0.3, 1.7, len(d))).astype(np.float32)
    centre[::7] = d[::7]
    io = np.zeros(len(rays), R.HITREC)
    io["prim"] = 0xffffffff
    io["dist"] = (2 * centre).astype(np.float32)
    rays["min_dist"] = 0.0
    return rays, io, centre


if __name__ == "__main__":
    build()
    pack = {}
    for name in GOLDEN_NAMES:
        g = Golden(name)
        ref = Ref(g.scene, threads=1).build()
        assert np.array_equal(ref.primid(), g.primid)        # same tree as the traversal fixture
        rays, io, centre = queries(g, 2500, 11)
        r, h = ref.closest(rays, io, centre)
        ref.close()
        for k, v in (("rays", rays), ("io", io), ("out_rays", r), ("out", h)):
            pack[f"{name}_{k}"] = v.view("u1").reshape(len(v), -1)
        pack[f"{name}_centre"] = centre
        print(name, "found", float((R.hit_prim64(h) != R.INVALID_PRIMID).mean()), "straight", float(np.mean(np.abs(h["dist"] - centre) <= 1e-6)))
    np.savez_compressed(os.path.join(HERE, "closest.npz"), **pack)
    print("wrote closest.npz", os.path.getsize(os.path.join(HERE, "closest.npz")) // 1024, "KiB")
