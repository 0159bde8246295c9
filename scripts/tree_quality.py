"""scratch: traversal work per ray (ACCEL_DEBUG-style counters of the oracle's CPU traversal) on the GPU-built LBVH tree vs the
reference-style binned-SAH tree (oracle builder) for the same scene and rays"""
import sys, os, time, importlib
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
cb = importlib.import_module("corona-13_b200")
lib = importlib.import_module("corona-13_b200.lib")
from oracle.binding import Oracle
S = cb.scenes
tris = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
sc = S.synthetic_scene(tris, seed=1)
n = 300_000
cam = S.camera_rays(n, sc)
rnd = S.random_rays(n, sc)
t = time.time(); acc = lib.Accel(sc).build(); print("gpu build", time.time() - t, "nodes", acc.num_nodes())
nodes, primid = acc.export_qbvh()
o1 = Oracle(sc).import_tree(nodes, acc.aabb(), primid)
t = time.time(); o2 = Oracle(sc).build(); print("oracle SAH build", time.time() - t, "nodes", len(o2.nodes()))
hits = None
for name, rays in (("camera", cam), ("random", rnd)):
    h1, c1 = o1.intersect(rays, counters=True)
    h2, c2 = o2.intersect(rays, counters=True)
    # secondary: diffuse-ish bounce from camera hits
    print(f"{name}: LBVH  boxes/ray {c1[1]/n:.2f} hit-boxes {c1[2]/n:.2f} prims/ray {c1[3]/n:.2f}")
    print(f"{name}: SAH   boxes/ray {c2[1]/n:.2f} hit-boxes {c2[2]/n:.2f} prims/ray {c2[3]/n:.2f}")
    if name == "camera": hits = h1
sr, md = S.shadow_rays(cam, hits, (0.0, 0.0, 9.0))
# bounce rays: from hit points in random directions
from importlib import import_module
R = cb.records
ok = R.hit_prim64(hits) != R.INVALID_PRIMID
b = cam[ok].copy()
rng = np.random.default_rng(9)
d = rng.normal(size=(len(b), 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
b["pos"] = cam[ok]["pos"] + hits[ok]["dist"][:, None] * cam[ok]["dir"] + 1e-3 * d
b["dir"] = d.astype(np.float32)
h1, c1 = o1.intersect(b, counters=True); h2, c2 = o2.intersect(b, counters=True)
m = len(b)
print(f"bounce: LBVH  boxes/ray {c1[1]/m:.2f} hit-boxes {c1[2]/m:.2f} prims/ray {c1[3]/m:.2f}")
print(f"bounce: SAH   boxes/ray {c2[1]/m:.2f} hit-boxes {c2[2]/m:.2f} prims/ray {c2[3]/m:.2f}")
