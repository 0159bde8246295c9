"""Generates the golden vectors in this directory from the UNMODIFIED reference compiled in place
(oracle/_ref/libcorona_ref.so, see oracle/Makefile and oracle/ref_glue.c).  Run in the build
container only (needs /root/reference):

    python tests/golden/make_golden.py

Each .npz holds one small scene (the raw .geo-layout arrays, so nothing depends on regenerating the
geometry bit-identically elsewhere), ray sets, and what the reference returned for them:
  hits_*      accel_intersect results (prim,u,v,dist), qbvhmp.c:1262
  vis         accel_visible results, qbvhmp.c:1392
  nodes, primid, aabb   the tree accel_build produced with one thread, qbvhmp.c:1181
  bounds0/1   prims_get_bounds_shutter_open/close per primitive, prims.c:20-60
"""
import importlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
cb = importlib.import_module("corona-13_b200")
from oracle.binding import Ref, build  # noqa: E402

S, R = cb.scenes, cb.records


def pack_scene(sc):
    d = {"num_shapes": np.int64(len(sc.shapes))}
    for i, s in enumerate(sc.shapes):
        d[f"s{i}_primid"] = s.primid
        d[f"s{i}_vtxidx"] = s.vtxidx.view("<u4").reshape(-1, 2)
        d[f"s{i}_vtx"] = s.vtx.view("<u4").reshape(-1, 4)
        d[f"s{i}_material"] = np.int64(s.material)
    return d


def edge_rays(sc):
    """axis-parallel rays (zero direction components -> inf/NaN slabs), rays starting on box planes,
    zero-length and far-away rays"""
    lo, hi = sc.bounds()
    ctr = (lo + hi) / 2
    pos, dr = [], []
    for ax in range(3):
        for sgn in (1.0, -1.0):
            d = np.zeros(3, np.float32)
            d[ax] = sgn
            for off in (0.0, 0.25, -0.25):
                p = ctr.copy()
                p[ax] = (lo[ax] - 1.0) if sgn > 0 else (hi[ax] + 1.0)
                p[(ax + 1) % 3] += off * (hi - lo)[(ax + 1) % 3]
                pos.append(p)
                dr.append(d)
    # origin exactly on the scene box planes, direction inside the plane
    for ax in range(3):
        p = ctr.copy()
        p[ax] = lo[ax]
        d = np.zeros(3, np.float32)
        d[(ax + 1) % 3] = 1.0
        pos.append(p)
        dr.append(d)
    pos.append(ctr + 1e6)
    dr.append(np.float32([0, 0, -1]))
    pos.append(ctr)
    dr.append(np.float32([-0.0, 0.0, -1.0]))
    return R.make_rays(np.asarray(pos, np.float32), np.asarray(dr, np.float32), 0.0)


def make(name, sc, time_max, n=1500):
    ref = Ref(sc, threads=1).build()
    prim = S.camera_rays(n, sc, seed=11, time_max=time_max)
    rnd = S.random_rays(n, sc, seed=12, time_max=time_max)
    edge = edge_rays(sc)
    rays = np.concatenate([prim, rnd, edge])
    hits = ref.intersect(rays, nthreads=1)
    bounce = S.bounce_rays(rays, hits, seed=13)
    hits_b = ref.intersect(bounce, nthreads=1)
    shadow, smd = S.shadow_rays(rays, hits, (0.0, 0.0, 9.0), seed=14)
    vis = ref.visible(shadow, smd, nthreads=1)
    # limited search distance (hit->dist preset below the first hit for half of the rays)
    md = np.where(np.arange(len(rays)) % 2 == 0, hits["dist"] * np.float32(0.5), hits["dist"]).astype(np.float32)
    md[~np.isfinite(md)] = R.FLT_MAX
    hits_md = ref.intersect(rays, md, nthreads=1)
    primid = ref.primid()
    b0 = np.stack([ref.prim_bounds(p, False) for p in primid])
    b1 = np.stack([ref.prim_bounds(p, True) for p in primid])
    out = pack_scene(sc)
    out.update(rays=rays.view("<u4").reshape(-1, 10), hits=hits.view("<u4").reshape(-1, 6),
               bounce=bounce.view("<u4").reshape(-1, 10), hits_bounce=hits_b.view("<u4").reshape(-1, 6),
               shadow=shadow.view("<u4").reshape(-1, 10), shadow_max_dist=smd, vis=vis,
               max_dist=md, hits_md=hits_md.view("<u4").reshape(-1, 6),
               nodes=ref.nodes().view("<u4").reshape(-1, 64), primid=primid, aabb=ref.aabb(),
               bounds0=b0, bounds1=b1)
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **out)
    print(name, sc.num_prims, "prims", len(rays), "rays", os.path.getsize(path) // 1024, "KiB",
          "hit rate %.2f" % (R.hit_prim64(hits) != R.INVALID_PRIMID).mean(), "vis %.2f" % vis.mean())
    ref.close()


if __name__ == "__main__":
    build()
    make("static_tris", S.synthetic_scene(1500, seed=21), 0.0)
    make("motion_quads_analytic", S.synthetic_scene(1200, seed=22, motion=True, quads=True, analytic=True), 1.0)
    geo = os.path.join(ROOT, "oracle", "_ref", "scenes", "geo")
    make("c10_geometry", S.c10_like_scene(geo if os.path.isdir(geo) else None), 0.24, n=1200)
