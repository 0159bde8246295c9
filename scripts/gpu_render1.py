"""scratch: GPU renders vs reference golden images, all cases and variants"""
import sys, os, time, importlib
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
from helpers import GoldenImage, image_stats, cb
lib = importlib.import_module("corona-13_b200.lib")
for case in (sys.argv[1:] or ["diffuse_static", "c10", "motion", "glass_metal"]):
    g = GoldenImage(case)
    acc = lib.Accel(g.scene).build()
    for key in g.variants:
        t = time.time()
        img, st = g.render(lib, acc, key)
        dt = time.time() - t
        a, b = g.ref(key, 1), g.ref(key, 2)
        noise, _ = image_stats(a, b)
        rel, ratio = image_stats(a, img)
        rel2, _ = image_stats(b, img)
        print(f"{case}/{key}: ref mean {a.mean(axis=(0,1))} gpu {img.mean(axis=(0,1))} ratio {ratio} relRMSE vs seed1 {rel:.4f} vs seed2 {rel2:.4f} "
              f"noise floor {noise:.4f} nan {np.isnan(img).sum()} time {dt:.2f}s {st}", flush=True)
    acc.close()
