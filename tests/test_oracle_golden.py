"""The oracle (oracle/oracle.c) against the golden vectors the unmodified reference produced
(tests/golden/make_golden.py): build, bounds, closest hit, limited-distance search, shadow rays."""
import numpy as np
import pytest

from helpers import Golden, GOLDEN_NAMES, R, assert_hits_equal, reachable_nodes
from oracle.binding import Oracle


@pytest.fixture(scope="module", params=GOLDEN_NAMES)
def g(request, built):
    return Golden(request.param)


def test_build_matches_reference_tree(g):
    """serial restatement of the binned-SAH build == reference with one thread: same node records, same
    primid permutation, same scene box (qbvhmp.c:875-1186)"""
    orc = Oracle(g.scene).build()
    assert np.array_equal(orc.primid(), g.primid)
    assert np.array_equal(orc.aabb().view("u4"), g.aabb.view("u4"))
    mine = orc.nodes()
    idx = reachable_nodes(g.nodes)
    assert np.array_equal(idx, reachable_nodes(mine))
    for f in ("aabb0", "aabb1", "child", "axis0", "axis00", "axis01"):
        assert np.array_equal(np.ascontiguousarray(mine[f][idx]).view("u1"), np.ascontiguousarray(g.nodes[f][idx]).view("u1")), f
    assert np.array_equal(mine["parent"][idx], g.nodes["parent"][idx])
    rc, stats = orc.check()
    assert rc == 0 and stats[3] == g.scene.num_prims
    orc.close()


def test_prim_bounds(g):
    orc = Oracle(g.scene).build()
    vc = R.primid_vcnt(g.primid)
    for i, p in enumerate(g.primid):
        b0, b1 = orc.prim_bounds(p, False), orc.prim_bounds(p, True)
        if vc[i] == R.PRIM_LINE:   # atan2f/sinf/cosf in the bounds: same libm here, but keep a tolerance
            assert np.allclose(b0, g.bounds0[i], atol=1e-5) and np.allclose(b1, g.bounds1[i], atol=1e-5)
        else:
            assert np.array_equal(b0.view("u4"), g.bounds0[i].view("u4"))
            assert np.array_equal(b1.view("u4"), g.bounds1[i].view("u4"))
    orc.close()


def test_closest_hit(g):
    orc = Oracle(g.scene).import_tree(g.nodes, g.aabb, g.primid)
    assert_hits_equal(orc.intersect(g.rays), g.hits, "primary/random/edge rays")
    assert_hits_equal(orc.intersect(g.bounce), g.hits_bounce, "bounce rays (ignore prim, offset origin)")
    assert_hits_equal(orc.intersect(g.rays, g.max_dist), g.hits_md, "preset hit->dist")
    orc.close()


def test_own_tree_gives_same_hits(g):
    """the restated builder's tree must give the reference's answers too"""
    orc = Oracle(g.scene).build()
    assert_hits_equal(orc.intersect(g.rays), g.hits, "own tree")
    orc.close()


def test_shadow_rays(g):
    orc = Oracle(g.scene).import_tree(g.nodes, g.aabb, g.primid)
    assert np.array_equal(orc.visible(g.shadow, g.shadow_max_dist), g.vis)
    orc.close()


def test_empty_scene(cb, built):
    """qbvhmp.c:1081-1099: root of four empty leaves, nothing is ever hit"""
    sc = cb.scenes.Scene([], "empty")
    orc = Oracle(sc).build()
    n = orc.nodes()
    assert len(n) == 1 and all(int(c) == 1 << 63 for c in n["child"][0])
    rays = cb.records.make_rays(np.zeros((4, 3), np.float32), np.float32([[0, 0, 1]] * 4))
    h = orc.intersect(rays)
    assert (R.hit_prim64(h) == R.INVALID_PRIMID).all() and (h["dist"] == R.FLT_MAX).all()
    assert (orc.visible(rays, np.full(4, 10.0, np.float32)) == 1).all()
    orc.close()


def test_closest_matches_reference(built):
    """accel_closest (qbvhmp.c:1493-1600): oracle restatement == the reference's recorded answers, incl. the mutated ray.min_dist"""
    import os
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "closest.npz"))
    for name in GOLDEN_NAMES:
        g = Golden(name)
        orc = Oracle(g.scene).import_tree(g.nodes, g.aabb, g.primid)
        rec = lambda k, dt: np.ascontiguousarray(z[f"{name}_{k}"]).view(dt).reshape(-1)
        r, h = orc.closest(rec("rays", R.RAY), rec("io", R.HITREC), z[f"{name}_centre"])
        assert_hits_equal(h, rec("out", R.HITREC), f"{name} closest")
        assert np.array_equal(r["min_dist"].view("u4"), rec("out_rays", R.RAY)["min_dist"].view("u4"))
        orc.close()
