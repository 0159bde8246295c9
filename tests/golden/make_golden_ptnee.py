"""Golden images for the `ptnee` sampler (src/sampler.d/ptnee.c: path tracing with next-event estimation only) from the
unmodified reference renderer (oracle/_ref/corona_ptnee_halton).  Run in the build container only:

    python tests/golden/make_golden_ptnee.py

Takes the scenes of existing image fixtures (tests/golden/img_<case>.npz), writes them out in the reference's own file formats
and stores the reference's renders for two --frame seeds in tests/golden/ptnee.npz (the second seed gives the noise floor).
Halton points with --frame 1: the GPU integrator draws the same samples.
"""
import os
import shutil
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import GoldenImage, cb   # noqa: E402

IO = cb.scene_io
REFDIR = os.path.join(ROOT, "oracle", "_ref")
CASES = ["diffuse_static", "glass_metal", "sky_light"]


def main():
    pack = {"cases": np.array(CASES)}
    for case in CASES:
        g = GoldenImage(case)
        tmp = tempfile.mkdtemp(prefix="corona_ptnee_")
        try:
            nra2 = g.write_files(tmp)
            for seed in (1, 2):
                cmd = [os.path.join(REFDIR, "corona_ptnee_halton"), nra2, "-x", "-s", str(g.spp), "-w", str(g.w), "-h", str(g.h), "-b", "0",
                       "-t", str(os.cpu_count()), "--frame", str(seed), "-q"]
                subprocess.run(cmd, cwd=REFDIR, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
                img = IO.read_pfm(os.path.splitext(nra2)[0] + "render_fb00.pfm")
                print(case, seed, img.shape, img.mean(axis=(0, 1)), "ptdl:", g.ref("ptdl_halton", 1).mean(axis=(0, 1)))
                pack[f"{case}_seed{seed}"] = img.astype(np.float32)
        finally:
            shutil.rmtree(tmp, ignore_errors=True)
    path = os.path.join(HERE, "ptnee.npz")
    np.savez_compressed(path, **pack)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
