"""The drop-in INSIDE the reference binary (run with -m gpu on the B200 box; the binaries are built where /root/reference exists,
oracle/Makefile target `intree`, and travel under oracle/_ref/).

    corona_b200_intree_<sampler>_<points>   the reference's main.c / shader.c / pathspace.c / prims.c / sampler / point sampler / lights /
                                            camera / display, its view.c with corona-13_b200/host/intree/view_render_b200.patch, and
                                            MOD_accel=b200 + MOD_render=b200 (host/accel_b200.c, render_b200.c, scene_b200.c) on
                                            libcorona_b200.so
    corona_accel_b200_pt_halton             MOD_accel=b200 alone under the reference's unmodified cpu renderer (render.d/gi.c): every
                                            accel_intersect of the worker becomes a one-ray call into the library

Checks: the in-tree binary, started exactly like the reference binary (same scene files, same command line), writes the reference's
artefacts -- <base>render_fb00.pfm, the sidecar .txt, the mmapped .fb -- and its image matches the reference renderer's
(tests/golden/img_*.npz) like the stand-alone paths do; MOD_accel=b200 alone reproduces the cpu renderer's image pixel for pixel
up to equal-distance ties.
"""
import os
import subprocess

import numpy as np
import pytest

from helpers import GoldenImage, image_stats, cb

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFDIR = os.path.join(ROOT, "oracle", "_ref")
TABLES = os.path.join(ROOT, "corona-13_b200", "data", "ref_tables.cbt")


def run_binary(name, nra2, *args, threads=None):
    exe = os.path.join(REFDIR, name)
    if not os.path.exists(exe):
        pytest.fail(f"{exe} is missing: `make -C oracle intree` where /root/reference exists (python -c 'import __graft_entry__ as g; g.build()')")
    cmd = [exe, nra2, "-x", "-b", "0", "-t", str(threads or os.cpu_count()), *args]
    env = dict(os.environ, CORONA_B200_TABLES=TABLES)
    return subprocess.run(cmd, cwd=REFDIR, capture_output=True, text=True, env=env, timeout=900)


@pytest.mark.parametrize("case,key", [("c10", "pt_halton"), ("c10", "ptdl_halton"), ("glass_metal", "ptdl_halton"), ("diffuse_static", "ptdl_rand")])
def test_reference_binary_with_b200_modules(built, tmp_path, case, key):
    IO = cb.scene_io
    g = GoldenImage(case)
    nra2 = g.write_files(str(tmp_path))
    p = run_binary("corona_b200_intree_" + key, nra2, "-s", str(g.spp), "-w", str(g.w), "-h", str(g.h), "--frame", "1", "--retain-framebuffer")
    assert p.returncode == 0, p.stderr[-2000:] + p.stdout[-2000:]
    assert "rendered" in p.stdout, p.stdout[-1500:]      # main.c's summary
    base = os.path.join(str(tmp_path), "test")
    img = IO.read_pfm(base + "render_fb00.pfm")
    a, b = g.ref(key, 1), g.ref(key, 2)
    assert img.shape == a.shape and np.isfinite(img).all()
    noise, _ = image_stats(a, b)
    rel, ratio = image_stats(a, img)
    if "halton" in key:
        assert rel <= 0.45 * noise, f"{case}/{key}: relRMSE {rel:.4f} vs noise floor {noise:.4f}"
        assert np.all(np.abs(ratio - 1) < 0.01), ratio
    else:
        rel2, _ = image_stats(b, img)
        assert min(rel, rel2) <= 1.15 * noise, f"{case}/{key}: relRMSE {rel:.4f}/{rel2:.4f} vs noise floor {noise:.4f}"
    # the reference's own artefacts, written by ITS code from the framebuffer the module filled: sidecar (corona_common.c:70-97,
    # view_print_info: samples per pixel) and the mmapped frame buffer file (framebuffer.h:76-140)
    side = open(base + "render_fb00.pfm.txt").read()
    assert f"samples per pixel: {g.spp}" in side and "accel    : b200" in side and "render   : b200" in side, side[:1200]
    fbfile = base + "_render_fb00.fb"
    assert os.path.exists(fbfile)
    hdr = np.fromfile(fbfile, np.uint64, 3)
    assert int(hdr[0]) == 1936686951 and (int(hdr[1]), int(hdr[2])) == (img.shape[1], img.shape[0])
    raw = np.fromfile(fbfile, np.float32, offset=32).reshape(img.shape)
    gain = np.fromfile(fbfile, np.float32, 1, offset=28)[0]
    assert np.allclose(raw * gain, img, rtol=1e-5, atol=1e-7)


@pytest.mark.parametrize("threads", [1, 4])
def test_mod_accel_b200_alone_under_the_cpu_renderer(built, tmp_path, threads):
    """an unpatched reference renderer on top of the GPU accel: one worker thread, one ray per call (the only thing an untouched
    checkout would do with MOD_accel=b200); with Halton points and one thread both binaries trace the same paths, so the images
    are equal except where a hit is an exact tie between two primitives (the trees differ).  Four threads: the workers call
    accel_intersect concurrently, each through its own stream and pinned slot (abi.cu: run_small); the same paths are summed in
    another order."""
    import time
    IO = cb.scene_io
    g = GoldenImage("diffuse_static")
    nra2 = g.write_files(str(tmp_path))
    args = ("-s", "2" if threads == 1 else "8", "-w", "64", "-h", "32", "--frame", "1")
    t0 = time.time()
    p = run_binary("corona_accel_b200_pt_halton", nra2, *args, threads=threads)
    dt = time.time() - t0
    assert p.returncode == 0, p.stderr[-2000:] + p.stdout[-2000:]
    assert "accel    : b200" in open(os.path.join(str(tmp_path), "testrender_fb00.pfm.txt")).read()
    img = IO.read_pfm(os.path.join(str(tmp_path), "testrender_fb00.pfm")).copy()
    q = run_binary("corona_pt_halton", nra2, *args, threads=threads)
    assert q.returncode == 0, q.stderr[-2000:]
    want = IO.read_pfm(os.path.join(str(tmp_path), "testrender_fb00.pfm"))
    assert img.shape == want.shape == (32, 64, 3)
    print(f"MOD_accel=b200 alone, {threads} thread(s): {dt:.2f} s for {args[1]} x 64 x 32 paths (process start, scene upload and build included)")
    if threads > 1:
        # which worker draws which path index -- and with it path->tangent_frame_scrambling, a points_rand() of the worker
        # (src/pathspace.c:212-213) -- depends on scheduling: two runs of the unmodified reference differ at noise level too
        assert np.isfinite(img).all() and abs(float(img.mean())/float(want.mean()) - 1.0) < 0.2, (float(img.mean()), float(want.mean()))
        return
    differ = np.abs(img - want).max(axis=2) > 1e-6 * max(float(want.max()), 1e-30)
    assert differ.mean() <= 0.01, f"{int(differ.sum())} of {differ.size} pixels differ between MOD_accel=b200 and qbvhmp under the same renderer"
