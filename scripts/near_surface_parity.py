"""scratch: rays that start a hair's breadth inside closed objects (dense media put volume vertices there): GPU vs oracle"""
import sys, os, importlib
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
from helpers import GoldenImage
from oracle.binding import Oracle
cb = importlib.import_module("corona-13_b200")
lib = importlib.import_module("corona-13_b200.lib")
R = cb.records
g = GoldenImage("skin")
acc = lib.Accel(g.scene).build()
nodes, primid = acc.export_qbvh()
orc = Oracle(g.scene).import_tree(nodes, acc.aabb(), primid)
rng = np.random.default_rng(3)
n = 200000
rays = np.zeros(n, R.RAY)
# half: inside the sphere (1.1,0.2,0.9) r .88 near its surface; half: inside the box (-1.8,-0.9,0.02)-(-0.2,0.7,1.6) near a face
d = rng.normal(size=(n, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
depth = 10 ** rng.uniform(-5, -2, n)
p = np.zeros((n, 3))
h = n // 2
u = rng.normal(size=(h, 3)); u /= np.linalg.norm(u, axis=1, keepdims=True)
p[:h] = np.array([1.1, 0.2, 0.9]) + u * (0.88 - depth[:h, None])
lo, hi = np.array([-1.8, -0.9, 0.02]), np.array([-0.2, 0.7, 1.6])
q = lo + rng.random((n - h, 3)) * (hi - lo)
ax = rng.integers(0, 3, n - h); side = rng.integers(0, 2, n - h)
q[np.arange(n - h), ax] = np.where(side == 1, hi[ax] - depth[h:], lo[ax] + depth[h:])
p[h:] = q
rays["pos"] = p.astype(np.float32); rays["dir"] = d.astype(np.float32)
rays["ignore"] = 0xffffffff
got = acc.intersect(rays); want = orc.intersect(rays)
print("intersect equal:", np.array_equal(got.view("u1"), want.view("u1")), "misses", int((R.hit_prim64(want) == R.INVALID_PRIMID).sum()))
md = np.full(n, 10.0, np.float32)
gv, wv = acc.visible(rays, md), orc.visible(rays, md)
print("visible equal:", np.array_equal(gv, wv), "visible count gpu/oracle", int(gv.sum()), int(wv.sum()))
# clipped closest hits like free-flight sampling does
clip = (10 ** rng.uniform(-4, -1, n)).astype(np.float32)
got = acc.intersect(rays, clip); want = orc.intersect(rays, clip)
print("clipped intersect equal:", np.array_equal(got.view("u1"), want.view("u1")), "surface hits", int((R.hit_prim64(want) != R.INVALID_PRIMID).sum()))
bad = np.nonzero((got.view("u1").reshape(n, -1) != want.view("u1").reshape(n, -1)).any(axis=1))[0]
print("mismatches", len(bad), got[bad[:5]], want[bad[:5]], clip[bad[:5]])
