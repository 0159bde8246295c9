"""shared test helpers: golden fixtures -> scenes / record arrays, comparison utilities"""
import importlib
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden")
cb = importlib.import_module("corona-13_b200")
S, R = cb.scenes, cb.records

GOLDEN_NAMES = ["static_tris", "motion_quads_analytic", "c10_geometry"]


class Golden:
    def __init__(self, name):
        z = np.load(os.path.join(GOLDEN, name + ".npz"))
        self.name = name
        shapes = []
        for i in range(int(z["num_shapes"])):
            shapes.append(S.Shape(z[f"s{i}_primid"], np.ascontiguousarray(z[f"s{i}_vtxidx"]).view(R.VTXIDX).reshape(-1),
                                  np.ascontiguousarray(z[f"s{i}_vtx"]).view(R.VTX).reshape(-1), int(z[f"s{i}_material"]), f"s{i}"))
        self.scene = S.Scene(shapes, name)
        rec = lambda k, dt: np.ascontiguousarray(z[k]).view(dt).reshape(-1)
        self.rays = rec("rays", R.RAY)
        self.hits = rec("hits", R.HITREC)
        self.bounce = rec("bounce", R.RAY)
        self.hits_bounce = rec("hits_bounce", R.HITREC)
        self.shadow = rec("shadow", R.RAY)
        self.shadow_max_dist = z["shadow_max_dist"]
        self.vis = z["vis"]
        self.max_dist = z["max_dist"]
        self.hits_md = rec("hits_md", R.HITREC)
        self.nodes = rec("nodes", R.QBVH_NODE)
        self.primid = z["primid"]
        self.aabb = z["aabb"]
        self.bounds0 = z["bounds0"]
        self.bounds1 = z["bounds1"]


def analytic_mask(hits):
    """hits on sphere / line prims: their u,v go through atan2f/acosf (libm, not bit-reproducible)"""
    vc = R.primid_vcnt(R.hit_prim64(hits))
    return (vc == R.PRIM_SPHERE) | (vc == R.PRIM_LINE)


UV_TOL = 2e-6   # absolute tolerance on u,v of analytic prims (values in [-0.5, 1])


def assert_hits_equal(got, want, what=""):
    """prim id and dist bit-exact everywhere; u,v bit-exact on triangles/quads, within UV_TOL on analytic prims"""
    gp, wp = R.hit_prim64(got), R.hit_prim64(want)
    bad = np.nonzero(gp != wp)[0]
    assert len(bad) == 0, f"{what}: {len(bad)} prim id mismatches, first at ray {bad[:5]}"
    bad = np.nonzero(got["dist"].view("u4") != want["dist"].view("u4"))[0]
    assert len(bad) == 0, f"{what}: {len(bad)} dist mismatches, first at ray {bad[:5]}"
    an = analytic_mask(want)
    hit = wp != R.INVALID_PRIMID
    for f in ("u", "v"):
        tri = hit & ~an
        bad = np.nonzero(got[f].view("u4")[tri] != want[f].view("u4")[tri])[0]
        assert len(bad) == 0, f"{what}: {len(bad)} {f} mismatches on triangles/quads"
        if an.any():
            err = np.abs(got[f][an] - want[f][an]).max()
            assert err <= UV_TOL, f"{what}: analytic {f} off by {err}"


def reachable_nodes(nodes):
    """indices of the nodes reachable from the root (the reference leaves abandoned nodes in its buffer)"""
    seen = []
    st = [0]
    while st:
        i = st.pop()
        seen.append(i)
        for c in nodes["child"][i]:
            if not (int(c) >> 63):
                st.append(int(c))
    return np.array(sorted(seen))


def classify_mismatches(orc, rays, got, want, max_dist=None):
    """mode-B bookkeeping: every ray where the GPU tree's answer differs from the reference tree's must be a
    proven tie -- the oracle's single-primitive test of the *other* primitive returns the identical distance
    bits.  returns (num_mismatch, num_proven_ties)"""
    gp, wp = R.hit_prim64(got), R.hit_prim64(want)
    idx = np.nonzero(gp != wp)[0]
    ties = 0
    for i in idx:
        if got["dist"][i].view("u4") != want["dist"][i].view("u4"):
            continue
        # same distance through a different primitive: test the GPU's prim alone with the oracle
        h = np.zeros(1, R.HIT)
        h["prim"] = 0xFFFFFFFF
        h["dist"] = R.FLT_MAX if max_dist is None else max_dist[i]
        orc.prim_intersect(gp[i], rays[i:i + 1], h)
        if h["dist"][0].view("u4") == want["dist"][i].view("u4"):
            ties += 1
    return len(idx), ties
