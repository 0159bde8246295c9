"""profiling driver: 10 M-tri scene, one primary + one bounce + one shadow wave, few launches (for ncu)"""
import sys, os, importlib
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
cb = importlib.import_module("corona-13_b200")
lib = importlib.import_module("corona-13_b200.lib")
S, R = cb.scenes, cb.records
import torch
nt = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
n = int(sys.argv[2]) if len(sys.argv) > 2 else (1 << 22)
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
sc = S.synthetic_scene(nt, seed=1)
acc = lib.Accel(sc).build()
cam = S.camera_rays(n, sc, seed=100)
hits = acc.intersect(cam)
bounce = S.bounce_rays(cam, hits, seed=200)
shadow, smd = S.shadow_rays(cam, hits, (0.0, 0.0, 9.0), seed=300)
st = torch.cuda.current_stream().cuda_stream
bufs = []
for rays, md in ((cam, None), (bounce, None), (shadow, smd)):
    d_r = torch.from_numpy(rays.view('u1').reshape(-1)).cuda()
    d_md = torch.from_numpy(md).cuda() if md is not None else None
    d_o = torch.zeros(len(rays)*24, dtype=torch.uint8, device='cuda')
    bufs.append((d_r, d_md, d_o, len(rays)))
torch.cuda.synchronize()
for _ in range(reps):
    for d_r, d_md, d_o, m in bufs:
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        if d_md is None: acc.intersect_dev(d_r.data_ptr(), 0, d_o.data_ptr(), m, st)
        else: acc.visible_dev(d_r.data_ptr(), d_md.data_ptr(), d_o.data_ptr(), m, st)
        e1.record(); torch.cuda.synchronize()
        print(f"rays {m} ms {e0.elapsed_time(e1):.3f} Grays/s {m/e0.elapsed_time(e1)/1e6:.3f}", flush=True)
