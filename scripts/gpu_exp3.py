"""scratch: timing of the traversal kernels on primary (tile-ordered / random), bounce, shadow waves"""
import sys, os, time, importlib
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
cb = importlib.import_module("corona-13_b200")
lib = importlib.import_module("corona-13_b200.lib")
S, R = cb.scenes, cb.records
import torch
nt = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
sc = S.synthetic_scene(nt, seed=1)
acc = lib.Accel(sc).build()
W, H = 2048, 2048
n = W*H
st = torch.cuda.current_stream().cuda_stream
def timeit(rays, md=None, reps=5, counted=True):
    m = len(rays)
    d_r = torch.from_numpy(rays.view('u1').reshape(-1)).cuda()
    d_md = torch.from_numpy(md).cuda() if md is not None else None
    d_o = torch.zeros(m*24, dtype=torch.uint8, device='cuda')
    f = (lambda: acc.intersect_dev(d_r.data_ptr(), 0, d_o.data_ptr(), m, st)) if md is None else (lambda: acc.visible_dev(d_r.data_ptr(), d_md.data_ptr(), d_o.data_ptr(), m, st))
    for _ in range(2): f()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(reps): f()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)/reps
    c = ""
    if counted and md is None:
        cnt = acc.intersect_counted(d_r.data_ptr(), 0, d_o.data_ptr(), m)
        c = str(np.round(cnt[1:]/cnt[0], 2))
    return f"{m/ms/1e6:.3f} Grays/s ({ms:.2f} ms) {c}"
cam_t = S.camera_rays(n, sc, seed=100, frame=(W, H))
cam_r = S.camera_rays(n, sc, seed=100)
hits = acc.intersect(cam_t)
bounce = S.bounce_rays(cam_t, hits, seed=200)
shadow, smd = S.shadow_rays(cam_t, hits, (0.0, 0.0, 9.0), seed=300)
print("threshold", os.environ.get("CB200_PRIM_THRESHOLD"), "tris", nt)
print(" primary tiled :", timeit(cam_t))
print(" primary random:", timeit(cam_r))
print(" bounce        :", timeit(bounce))
print(" shadow        :", timeit(shadow, smd))
