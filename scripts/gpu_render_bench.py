"""scratch: one 4K ptdl pass on the 10M-triangle scene, per-kernel-class timing"""
import sys, os, time, importlib
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
cb = importlib.import_module("corona-13_b200")
lib = importlib.import_module("corona-13_b200.lib")
IO, S = cb.scene_io, cb.scenes
import ctypes as C
tris = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
z = np.load(os.path.join(ROOT, "corona-13_b200/data/bench_materials.npz"))
ms = IO.MaterialSet()
raw = z["materials"].tobytes()
arr = (IO.CMaterial * (len(raw)//C.sizeof(IO.CMaterial))).from_buffer_copy(raw)
ms.materials = list(arr)
sc = S.synthetic_scene(tris, seed=1, motion=bool(os.environ.get('BENCH_MOTION')), analytic=bool(os.environ.get('BENCH_ANALYTIC')))
mats = list(z['shape_mats']) + [7]*8
for s, m in zip(sc.shapes, mats): s.material = int(m)
acc = lib.Accel(sc).build()
cam = IO.Camera(pos=(18.0, 14.0, 14.5), lookat=(0.0, 0.0, 2.5), aperture_value=6, exposure_value=13, focal_length=0.4, iso=100.0)
for batch in [int(x) for x in os.environ.get('BENCH_BATCHES', '0').split(',')]:
    r = lib.Render(acc, cam, ms, 3840, 2176, sampler=1, pointsampler=0, frame=1, batch_paths=batch)
    if not os.environ.get('BENCH_NO_WARM'):   # (ncu captures: BENCH_NO_WARM=1, so that the first launches already are full streamed waves)
        r.render_pass()
        r.clear()
    r.instrument(True, False)
    t = time.time()
    NP = 8
    for k in range(NP):
        if k == 5 and len(sys.argv) > 2: os.environ['CB200_RENDER_TRACE'] = '1'
        r.render_pass(streaming=True)
        os.environ.pop('CB200_RENDER_TRACE', None)
    r.flush()
    dt = (time.time() - t)/NP
    st = r.stats()
    rays = (st["rays_closest"] + st["rays_shadow"])/NP
    print(f"batch {batch}: {dt*1e3:.1f} ms/pass, {rays/dt/1e6:.1f} Mrays/s, rays/path {rays/(3840*2176):.2f}, ms per class (all passes) {[round(x,1) for x in st['ms']]}, launches {st['kernel_launches']/NP:.0f}")
    img = r.image()
    print("  image mean", img.mean(axis=(0,1)), "nan", np.isnan(img).sum())
    r.close()
