"""scratch: white furnace.  Index-matched objects filled with a dense, non-absorbing medium under a constant sky and no other
light: no pixel may be brighter than the sky seen directly (the 32-vertex cap only removes energy)."""
import sys, os, importlib, tempfile
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
from helpers import GoldenImage
cb = importlib.import_module("corona-13_b200")
lib = importlib.import_module("corona-13_b200.lib")
IO, S = cb.scene_io, cb.scenes
g = GoldenImage("skin")
lines = [str(x) for x in g.z["shader_lines"]]
dens = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
lines = [l.replace("diffdiel 1.33 30", "dielectric 1.0 0").replace("color v 0.99 0.91 0.85", "color v 1 1 1")
          .replace("medium_rgb 0.014 0.005 0.003", f"medium_rgb {0.014*dens} {0.005*dens} {0.003*dens}") for l in lines]
shapes = g.scene.shapes[2:]          # cube and ball only: no emitter, no plane
tmp = tempfile.mkdtemp()
nra2 = os.path.join(tmp, "f.nra2")
IO.write_nra2(nra2, lines, [(12, "a"), (12, "b")], sky="sky_const 1 1 1 1")
r2s = IO.Rgb2Spec(IO.coeff_path(ROOT))
ms, _, _ = IO.parse_nra2(nra2, r2s)
for s in shapes: s.material = 12
sc = S.Scene(shapes, "furnace")
acc = lib.Accel(sc).build()
co, scale = IO.sky_const_params(r2s, "1 1 1 1")
for key in ("pt_halton", "ptdl_halton"):
    r = lib.Render(acc, g.camera, ms, 192, 128, frame=1, sky=IO.SKY_CONST, sky_coeff=co, sky_scale=scale, **GoldenImage.variant_args(key))
    for _ in range(64): r.render_pass()
    img = r.image()
    sky = np.median(img[:8, :8], axis=(0, 1))
    obj = img[40:100, 60:150]
    print(key, "sky", sky, "object mean/sky", obj.mean(axis=(0, 1)) / sky, "object max/sky", obj.max(axis=(0, 1)) / sky, r.stats()["rays_closest"] / r.stats()["paths"])
    r.close()
