"""scratch: per-kernel-class CUDA-event time of regression/0010_pt's geometry at 1024x576 (the named configs), streamed calls of
`batch` progressions each.  usage: named_classes.py [sampler=ptdl] [batch=16] [calls=16]"""
import sys, os, time, importlib
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import GoldenImage
lib = importlib.import_module("corona-13_b200.lib")
sampler = sys.argv[1] if len(sys.argv) > 1 else "ptdl"
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 16
calls = int(sys.argv[3]) if len(sys.argv) > 3 else 16
g = GoldenImage("c10")
acc = lib.Accel(g.scene).build()
W, H = 1024, 576
r = lib.Render(acc, g.camera, g.materials, W, H, frame=1, **g.sky_args, **GoldenImage.variant_args(sampler + "_rand"))
r.render_pass(0, W*H*batch, streaming=True)
r.flush(); r.clear()
r.instrument(True, False)
t = time.time()
for k in range(calls):
    if k == calls//2 and os.environ.get("TRACE"): os.environ['CB200_RENDER_TRACE'] = '1'
    r.render_pass(k*W*H*batch, W*H*batch, streaming=True)
    os.environ.pop('CB200_RENDER_TRACE', None)
r.flush()
dt = (time.time() - t)/(calls*batch)
st = r.stats()
print(f"{sampler} batch {batch}: {dt*1e3:.3f} ms/progression = {1/dt:.0f} spp/s; per progression, ms per class [start, closest, shade, shadow, resolve] "
      f"{[round(x/(calls*batch), 3) for x in st['ms']]}, sum {sum(st['ms'])/(calls*batch):.3f}; rays/path {(st['rays_closest'] + st['rays_shadow'])/st['paths']:.2f}, "
      f"closest {st['rays_closest']/st['ms'][1]/1e6:.2f} G rays/s, shadow {st['rays_shadow']/max(st['ms'][3], 1e-9)/1e6:.2f} G rays/s, launches per call {st['kernel_launches']/calls:.0f}")
r.close(); acc.close()
