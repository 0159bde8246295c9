"""Golden vectors for `--dbor n` (the density based outlier rejection cascade of view_splat_col, src/view.c:497-522) from the
unmodified reference renderer.  Run in the build container only:

    python tests/golden/make_golden_dbor.py

Takes the scene of an existing image fixture (tests/golden/img_<case>.npz), writes it out in the reference's own file
formats, runs oracle/_ref/corona_<variant> with `--dbor LEVELS` (two --frame seeds) and stores the framebuffer plus the LEVELS cascade images
(`<basename>render_dbor%02d.pfm`, view.c:553-556, scaled by the framebuffer's gain like the reference exports them) in
tests/golden/dbor.npz.  The variants use the Halton point sampler with --frame 1, so the GPU integrator draws the SAME
samples and every cascade level can be compared image against image.
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
RUNS = [("glass_metal", "ptdl_halton", 8), ("c10", "pt_halton", 12)]   # (image fixture, reference binary, --dbor levels)


def main():
    pack = {"runs": np.array([f"{c}:{v}:{n}" for c, v, n in RUNS])}
    for case, variant, levels in RUNS:
        g = GoldenImage(case)
        tmp = tempfile.mkdtemp(prefix="corona_dbor_")
        try:
            nra2 = g.write_files(tmp)
            for seed in (1, 2):   # the second seed gives the Monte Carlo noise floor of every level
                cmd = [os.path.join(REFDIR, "corona_" + variant), nra2, "-x", "-s", str(g.spp), "-w", str(g.w), "-h", str(g.h), "-b", "0",
                       "-t", str(os.cpu_count()), "--frame", str(seed), "--dbor", str(levels), "-q"]
                subprocess.run(cmd, cwd=REFDIR, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
                base = os.path.splitext(nra2)[0]
                fb = IO.read_pfm(base + "render_fb00.pfm")
                lv = np.stack([IO.read_pfm(base + f"render_dbor{l:02d}.pfm") for l in range(levels)])
                print(case, variant, seed, "fb mean", fb.mean(), "level means", lv.mean(axis=(1, 2, 3)))
                pack[f"{case}_{variant}_fb_seed{seed}"] = fb.astype(np.float32)
                pack[f"{case}_{variant}_dbor_seed{seed}"] = lv.astype(np.float32)
        finally:
            shutil.rmtree(tmp, ignore_errors=True)
    path = os.path.join(HERE, "dbor.npz")
    np.savez_compressed(path, **pack)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
