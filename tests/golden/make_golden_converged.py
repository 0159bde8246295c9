"""Long reference renders of the BASELINE configs AT THEIR STATED SIZES, for the reference's own regression criterion.

    python tests/golden/make_golden_converged.py [case ...]          (build container only: runs oracle/_ref/corona_*)

The reference judges a render by `pfmdiff testrender.pfm reference.pfm` <= regression/<case>/maxerror, reference.pfm being a
long render that is not available offline (regression/createres.sh:20-30, tools/img/pfmdiff.c:75-86).  This script makes those
long renders with the unmodified reference renderer on the scenes of img_c10.npz / img_motion.npz:

    c10_pt     regression/0010_pt   pt,   rand,   1024x576   (args: -s 128 -w 1024 -h 576, maxerror 4.0)
    c10_ptdl   regression/0011_ptdl ptdl, rand,   1024x576   (maxerror 3.8)
    motion     regression/0002_mb   ptdl, halton, rec709, 512x288 (maxerror 0.11)

and stores, per case, in tests/golden/conv_<case>.npz:
    long       the LONG-spp image as float16 (relative rounding 5e-4, far below the noise of a 128-spp test render)
    long_spp, long_frame
    ref_rmse   pfmdiff RMSE of the reference's OWN renders at 32 / 128 / 512 spp (other --frame) against `long`: what the
               criterion and the 1/sqrt(spp) convergence look like for the reference itself on this scene
The scene itself (geometry, materials, camera) is the one in img_c10.npz / img_motion.npz.
"""
import importlib
import os
import shutil
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import GoldenImage                      # noqa: E402
from make_golden_images import run_reference        # noqa: E402

CASES = {   # name: (image fixture, reference binary, w, h, long spp, maxerror)
    "c10_pt":   ("c10", "pt_rand", 1024, 576, 4096, 4.0),
    "c10_ptdl": ("c10", "ptdl_rand", 1024, 576, 4096, 3.8),
    "motion":   ("motion", "ptdl_halton_rec709", 512, 288, 8192, 0.11),
}
LONG_FRAME = 7
TEST_SPP = (32, 128, 512)


def pfmdiff_rmse(a, b):
    """tools/img/pfmdiff.c:75-86: sqrt(sum over pixels of the squared channel differences / (w*h))"""
    d = a.astype(np.float64) - b.astype(np.float64)
    return float(np.sqrt((d * d).sum() / (a.shape[0] * a.shape[1])))


def make(name):
    fixture, key, w, h, long_spp, maxerror = CASES[name]
    g = GoldenImage(fixture)
    tmp = tempfile.mkdtemp(prefix="corona_conv_")
    try:
        nra2 = g.write_files(tmp)
        long = run_reference("corona_" + key, nra2, w, h, long_spp, LONG_FRAME)
        print(name, "long", long.shape, long.mean(axis=(0, 1)), "max", long.max(), flush=True)
        rm = {}
        for spp in TEST_SPP:
            for frame in (1, 2):
                img = run_reference("corona_" + key, nra2, w, h, spp, frame)
                rm[(spp, frame)] = pfmdiff_rmse(img, long)
                print(name, "ref", spp, frame, "rmse", rm[(spp, frame)], flush=True)
        l16 = np.minimum(long, 65504.0).astype(np.float16)
        out = os.path.join(HERE, f"conv_{name}.npz")
        np.savez_compressed(out, long=l16, long_spp=np.int64(long_spp), long_frame=np.int64(LONG_FRAME), w=np.int64(w), h=np.int64(h),
                            key=np.array(key), fixture=np.array(fixture), maxerror=np.float64(maxerror),
                            clipped=np.int64((long > 65504.0).sum()),
                            ref_spp=np.int64(TEST_SPP), ref_rmse=np.float64([[rm[(s, f)] for f in (1, 2)] for s in TEST_SPP]))
        print("wrote", out, os.path.getsize(out) // 1024, "KiB", flush=True)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    for n in (sys.argv[1:] or list(CASES)):
        make(n)
