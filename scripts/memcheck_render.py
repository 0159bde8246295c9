"""small renders that touch every integrator code path added late (dbor cascade, ptnee, sphere lights, media, envmap), meant to be
run under `compute-sanitizer --tool memcheck` on the GPU box:  compute-sanitizer --tool memcheck python scripts/memcheck_render.py"""
import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import GoldenImage
lib = importlib.import_module("corona-13_b200.lib")
lib.set_device(0)
for case, key, dbor in [("sphere_light", "ptdl_halton", 6), ("sphere_light", "ptnee_halton", 0), ("glass_metal", "pt_halton", 4), ("fog", "ptdl_halton", 0),
                        ("envmap", "ptdl_halton", 3), ("motion", "ptdl_halton_rec709", 0), ("vstack", "ptnee_halton", 2)]:
    g = GoldenImage(case)
    acc = lib.Accel(g.scene).build()
    r = lib.Render(acc, g.camera, g.materials, g.w, g.h, frame=1, **g.sky_args, **GoldenImage.variant_args(key))
    if dbor:
        r.set_dbor(dbor)
    for _ in range(3):
        r.render_pass(streaming=True)
    img = r.image()
    lv = r.dbor_images() if dbor else np.zeros(1)
    print(case, key, dbor, "mean", float(img.mean()), "levels", float(lv.sum()), "finite", bool(np.isfinite(img).all()), flush=True)
    r.close(); acc.close()
print("done")
