"""scratch: first GPU render vs reference golden images"""
import sys, os, time, importlib
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
from helpers import GoldenImage, image_stats, cb
lib = importlib.import_module("corona-13_b200.lib")
IO = cb.scene_io
g = GoldenImage("diffuse_static")
acc = lib.Accel(g.scene).build()
for key, sampler, points in [("pt", 0, 0), ("ptdl", 1, 0), ("ptdl_halton", 1, 1)]:
    r = lib.Render(acc, g.camera, g.materials, g.w, g.h, sampler=sampler, pointsampler=points, frame=1)
    t = time.time()
    for _ in range(g.spp): r.render_pass()
    img = r.image()
    dt = time.time() - t
    a, b = g.ref(key, 1), g.ref(key, 2)
    noise, mr = image_stats(a, b)
    rel, ratio = image_stats(a, img)
    print(f"{key}: ref mean {a.mean(axis=(0,1))} gpu mean {img.mean(axis=(0,1))} ratio {ratio} relRMSE gpu-vs-ref {rel:.4f} noise floor (ref seeds) {noise:.4f} time {dt:.2f}s stats {r.stats()}", flush=True)
    r.close()
