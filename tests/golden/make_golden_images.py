"""Golden IMAGES from the unmodified reference renderer (oracle/_ref/corona_*, built by oracle/Makefile from the
reference sources in place).  Run in the build container only:

    python tests/golden/make_golden_images.py

For each case: writes the scene (.geo / .nra2 / .cam) to a scratch directory, runs the reference with two different
--frame seeds, and stores both images (fb*gain as the reference exports them) plus the scene description in
tests/golden/img_<case>.npz.  The second seed gives the Monte Carlo noise floor the GPU image is judged against.
"""
import importlib
import os
import shutil
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
cb = importlib.import_module("corona-13_b200")
S, IO = cb.scenes, cb.scene_io
REFDIR = os.path.join(ROOT, "oracle", "_ref")


def run_reference(binary, scene_dir, nra2, w, h, spp, frame, threads=None):
    """corona <scene> -x -s spp -w w -h h -b 0 --frame f, cwd = oracle/_ref (data/ergb2spec.coeff is cwd-relative, main.c:292)"""
    threads = threads or os.cpu_count()
    cmd = [os.path.join(REFDIR, binary), nra2, "-x", "-s", str(spp), "-w", str(w), "-h", str(h), "-b", "0",
           "-t", str(threads), "--frame", str(frame), "-q"]
    subprocess.run(cmd, cwd=REFDIR, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    base = os.path.splitext(nra2)[0]
    img = IO.read_pfm(base + "render_fb00.pfm") if os.path.exists(base + "render_fb00.pfm") else None
    if img is None:
        cand = [f for f in os.listdir(os.path.dirname(nra2)) if f.endswith(".pfm")]
        img = IO.read_pfm(os.path.join(os.path.dirname(nra2), cand[0]))
    return img


def synthetic_case(name, scene, shader_lines, shape_mats, cam, w, h, spp, binaries):
    tmp = tempfile.mkdtemp(prefix="corona_golden_")
    try:
        shapes = []
        for i, sh in enumerate(scene.shapes):
            sh.write_geo(os.path.join(tmp, f"shape{i}.geo"))
            shapes.append((shape_mats[i], f"shape{i}"))
        nra2 = os.path.join(tmp, "test.nra2")
        IO.write_nra2(nra2, shader_lines, shapes)
        cam.write(os.path.join(tmp, "test01.cam"))
        out = {}
        for key, binary in binaries.items():
            for seed in (1, 2):
                out[f"{key}_seed{seed}"] = run_reference(binary, tmp, nra2, w, h, spp, seed)
                print(name, key, seed, out[f"{key}_seed{seed}"].shape, out[f"{key}_seed{seed}"].mean(axis=(0, 1)))
        ms, _, _ = IO.parse_nra2(nra2, IO.Rgb2Spec(IO.coeff_path(ROOT)))
        mats, _ = ms.carrays()
        matbytes = np.frombuffer(bytes(mats), np.uint8)   # flattened cb_material_t[] incl. the rgb2spec coefficients
        pack = {"materials": matbytes, "num_shapes": np.int64(len(scene.shapes)), "shape_mats": np.int64(shape_mats), "w": np.int64(w), "h": np.int64(h),
                "spp": np.int64(spp), "shader_lines": np.array(shader_lines), "cam": np.frombuffer(open(os.path.join(tmp, "test01.cam"), "rb").read(), np.uint8)}
        for i, s in enumerate(scene.shapes):
            pack[f"s{i}_primid"] = s.primid
            pack[f"s{i}_vtxidx"] = s.vtxidx.view("<u4").reshape(-1, 2)
            pack[f"s{i}_vtx"] = s.vtx.view("<u4").reshape(-1, 4)
        for k, v in out.items():
            pack[k] = v.astype(np.float32)
        path = os.path.join(HERE, f"img_{name}.npz")
        np.savez_compressed(path, **pack)
        print("wrote", path, os.path.getsize(path) // 1024, "KiB")
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "ref"])
    # case 1: diffuse terrain + soup under a quad light, static
    sc = S.synthetic_scene(3000, seed=7)
    lines = ["diffuse", "color d 0.6 0.5 0.4", "mult 1 1 0", "color d 0 0 0", "color e 30 30 30 1.", "mult 2 3 4 0"]
    mats = [2, 2, 5]   # terrain, soup, light
    cam = IO.Camera(pos=(14.0, 11.0, 9.0), lookat=(0.0, 0.0, 2.0), aperture_value=6, exposure_value=13, focal_length=0.35, iso=400.0)
    synthetic_case("diffuse_static", sc, lines, mats, cam, 160, 96, 256,
                   {"pt": "corona_pt_rand", "ptdl": "corona_ptdl_rand", "ptdl_halton": "corona_ptdl_halton"})
