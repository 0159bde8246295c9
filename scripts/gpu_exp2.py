"""scratch: (1) debug the mode-B non-tie mismatch, (2) how much does tree quality matter (reference SAH tree imported)"""
import sys, os, time, importlib
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
cb = importlib.import_module("corona-13_b200")
lib = importlib.import_module("corona-13_b200.lib")
from oracle.binding import Oracle, Ref, ref_available
S, R = cb.scenes, cb.records
import torch

cfg = dict(num_tris=50000, seed=42, quads=True, motion=True)
sc = S.synthetic_scene(**cfg)
orc = Oracle(sc).build()
rays = np.concatenate([S.camera_rays(150000, sc, time_max=1.0), S.random_rays(150000, sc, time_max=1.0)])
want = orc.intersect(rays)
br = S.bounce_rays(rays, want)
want_b = orc.intersect(br)
acc = lib.Accel(sc).build()
nodes, primid = acc.export_qbvh()
chk = Oracle(sc).import_tree(nodes, acc.aabb(), primid)
got_b = acc.intersect(br)
gp, wp = R.hit_prim64(got_b), R.hit_prim64(want_b)
for i in np.nonzero(gp != wp)[0]:
    print("ray", i, br[i])
    print(" gpu ", hex(gp[i]), got_b[i])
    print(" ref ", hex(wp[i]), want_b[i])
    for nm, p in (("gpu prim", gp[i]), ("ref prim", wp[i])):
        if p == R.INVALID_PRIMID: continue
        h = np.zeros(1, R.HIT); h["prim"] = 0xFFFFFFFF; h["dist"] = R.FLT_MAX
        orc.prim_intersect(p, br[i:i+1], h)
        print("  oracle single-prim test of", nm, "-> dist", h["dist"][0], h["dist"].view('u4')[0], "u,v", h["u"][0], h["v"][0])
acc.close()

# tree quality experiment
def timeit(acc, rays, md=None, reps=5):
    n = len(rays)
    d_r = torch.from_numpy(rays.view('u1').reshape(-1)).cuda()
    d_o = torch.zeros(n*24, dtype=torch.uint8, device='cuda')
    st = torch.cuda.current_stream().cuda_stream
    for _ in range(3): acc.intersect_dev(d_r.data_ptr(), 0, d_o.data_ptr(), n, st)
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(reps): acc.intersect_dev(d_r.data_ptr(), 0, d_o.data_ptr(), n, st)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)/reps
    cnt = acc.intersect_counted(d_r.data_ptr(), 0, d_o.data_ptr(), n)
    return n/ms/1e6, cnt[1:]/cnt[0]

for nt in [1000000, 10000000]:
    sc = S.synthetic_scene(nt, seed=1)
    n = 1 << 22
    cam = S.camera_rays(n, sc, seed=100)
    acc = lib.Accel(sc).build()
    hits = acc.intersect(cam)
    bounce = S.bounce_rays(cam, hits, seed=200)
    print(nt, "LBVH  nodes", acc.num_nodes(), "camera", timeit(acc, cam), "bounce", timeit(acc, bounce), flush=True)
    t = time.time(); ref = Ref(sc, threads=os.cpu_count()).build(); print("  ref build s", time.time()-t)
    acc.import_qbvh(ref.nodes(), ref.primid(), ref.aabb())
    print(nt, "refSAH(256B nodes) nodes", acc.num_nodes(), "depth", acc.depth(), "camera", timeit(acc, cam), "bounce", timeit(acc, bounce), flush=True)
    acc.close(); ref.close()
