"""Pins the oracle against the live compiled reference (oracle/_ref/libcorona_ref.so) on more and larger
seeded scenes than the committed golden vectors cover.  Skipped where oracle/_ref has not been built
(it can only be built where /root/reference exists; the .so travels to the GPU box)."""
import numpy as np
import pytest

from helpers import S, R, assert_hits_equal, reachable_nodes
from oracle.binding import Oracle, Ref, ref_available

pytestmark = pytest.mark.skipif(not ref_available(), reason="oracle/_ref not built")

CASES = {
    "tris_5k": dict(num_tris=5000, seed=31),
    "quads_mb_8k": dict(num_tris=8000, seed=32, quads=True, motion=True),
    "analytic_mb_3k": dict(num_tris=3000, seed=33, analytic=True, motion=True),
    "tris_60k": dict(num_tris=60000, seed=34, analytic=True),
}


@pytest.fixture(scope="module", params=list(CASES))
def pair(request, built):
    cfg = CASES[request.param]
    sc = S.synthetic_scene(**cfg)
    ref = Ref(sc, threads=1).build()
    orc = Oracle(sc).build()
    yield sc, ref, orc, (1.0 if cfg.get("motion") else 0.0)
    ref.close()
    orc.close()


def test_build(pair):
    sc, ref, orc, _ = pair
    assert np.array_equal(ref.primid(), orc.primid())
    rn, on = ref.nodes(), orc.nodes()
    idx = reachable_nodes(rn)
    for f in ("aabb0", "aabb1", "child", "axis0", "axis00", "axis01"):
        assert np.array_equal(np.ascontiguousarray(rn[f][idx]).view("u1"), np.ascontiguousarray(on[f][idx]).view("u1")), f
    assert np.array_equal(ref.aabb().view("u4"), orc.aabb().view("u4"))


def test_traversal(pair):
    sc, ref, orc, tmax = pair
    rays = np.concatenate([S.camera_rays(20000, sc, time_max=tmax), S.random_rays(20000, sc, time_max=tmax)])
    want = ref.intersect(rays)
    assert_hits_equal(orc.intersect(rays), want, "closest")
    br = S.bounce_rays(rays, want)
    assert_hits_equal(orc.intersect(br), ref.intersect(br), "bounce")
    sr, md = S.shadow_rays(rays, want, (0.0, 0.0, 9.0))
    assert np.array_equal(orc.visible(sr, md), ref.visible(sr, md))


def test_counters_match_accel_debug(built):
    """the oracle's counters follow the reference's ACCEL_DEBUG definitions (qbvhmp.c:83-90,1307-1308,1375)"""
    if not ref_available(dbg=True):
        pytest.skip("debug build of the reference missing")
    sc = S.synthetic_scene(5000, seed=35)
    ref = Ref(sc, threads=1, dbg=True).build()
    orc = Oracle(sc).build()
    rays = S.camera_rays(5000, sc)
    ref.counters(reset=True)
    ref.intersect(rays, nthreads=1)
    rc = ref.counters()
    _, oc = orc.intersect(rays, counters=True)
    # reference order: {accel_intersect, aabb_intersect, aabb_true, prims_intersect}
    assert list(rc) == list(oc)
    ref.close()
    orc.close()
