"""Known-answer vectors for homogeneous media from the unmodified reference: medium_rgb (src/shaders/medium_rgb.c) behind an
optional `color v` (src/shaders/color.c, texture.h:46-52), the free-flight sampling / transmittance / distance pdf of
src/shader.c:46-131 and the phase-function callbacks at a volume vertex -- driven by oracle/ref_bsdf.c:ref_medium_eval.
Build container only:

    python tests/golden/make_golden_medium.py      ->  tests/golden/medium.npz

Stored per case: the queries, the reference's answers and the flattened cb_medium_t the product's own .nra2 reader
(scene_io.parse_nra2) produces for the same two shader lines, so the GPU test checks reader + kernels against the reference.
"""
import ctypes as C
import importlib
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
cb = importlib.import_module("corona-13_b200")
IO = cb.scene_io
REFDIR = os.path.join(ROOT, "oracle", "_ref")

# name: (medium_rgb arguments, `color v` arguments or None)
CASES = {"fog": ("9 12 16 0.6", "0.95 0.95 0.9"),
         "milk": ("0.5 0.35 0.25 0.0", "0.98 0.96 0.9"),
         "skin": ("0.3 0.12 0.08 0.7", "0.95 0.8 0.7"),
         "back": ("2 2.5 3 -0.5", "0.5 0.6 0.7"),
         "ink": ("4 1.5 0.8 0.0", None),                      # no albedo step: absorbs only (mu_s is NaN upstream)
         "thin": ("1e10 0.25 0.25 0.3", "0 0 0")}             # regression/0090_vstack style: black albedo


def queries(rng, n):
    q = np.zeros(n, IO.MEDIUM_QUERY)
    for key in ("wi", "wo"):
        d = rng.normal(size=(n, 3))
        q[key] = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    q["wi"][:6] = np.float32([[1, 0, 0], [0, 1, 0], [0, 0, 1], [-1, 0, 0], [0, -1, 0], [0, 0, -1]])   # onb branch points
    q["lambda_"] = rng.uniform(360.0, 830.0, n).astype(np.float32)
    q["rand"] = rng.random((n, 3), dtype=np.float32)
    q["rand"][:4, 2] = np.float32([0.0, 1e-7, 0.5, 0.99999994])
    q["dist"] = rng.exponential(4.0, n).astype(np.float32)
    return q


def flattened(medium_args, albedo_args):
    """the cb_medium_t scene_io.parse_nra2 makes of the two lines"""
    lines = ["diffuse"]
    if albedo_args is not None:
        lines += ["medium_rgb " + medium_args, "color v " + albedo_args, "mult 1 2 1", "interior 0 3"]
    else:
        lines += ["medium_rgb " + medium_args, "interior 0 1"]
    with tempfile.TemporaryDirectory() as tmp:
        path = os.path.join(tmp, "m.nra2")
        IO.write_nra2(path, lines, [(len(lines) - 1, "none")])
        ms, _, _ = IO.parse_nra2(path, IO.Rgb2Spec(IO.coeff_path(ROOT)))
    assert len(ms.media) == 1 and ms.materials[-1].medium == 1
    return np.frombuffer(bytes(ms.cmedia()), np.uint8).copy()


if __name__ == "__main__":
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "ref"])
    L = C.CDLL(os.path.join(REFDIR, "libref_bsdf.so"), mode=C.RTLD_GLOBAL)
    L.ref_bsdf_open.argtypes = [C.c_char_p, C.c_char_p]
    L.ref_medium_setup.argtypes = [C.c_char_p]
    L.ref_medium_eval.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_uint64]
    assert L.ref_medium_setup(IO.coeff_path(ROOT).encode()) == 0
    rng = np.random.default_rng(77)
    pack = {"cases": np.array(list(CASES))}
    for name, (margs, aargs) in CASES.items():
        hm = L.ref_bsdf_open(os.path.join(REFDIR, "shaders", "libmedium_rgb.so").encode(), margs.encode())
        ha = L.ref_bsdf_open(os.path.join(REFDIR, "shaders", "libcolor.so").encode(), ("v " + aargs).encode()) if aargs is not None else -1
        assert hm >= 0 and (ha >= 0 or aargs is None), name
        q = queries(rng, 4000)
        out = np.zeros(len(q), IO.MEDIUM_RESULT)
        L.ref_medium_eval(hm, ha, q.ctypes.data, out.ctypes.data, len(q))
        pack[name + "_q"], pack[name + "_r"] = q.view("u1").reshape(len(q), -1), out.view("u1").reshape(len(q), -1)
        pack[name + "_medium"] = flattened(margs, aargs)
        print(f"{name}: mu_t {out['mu_t'].min():.4g}..{out['mu_t'].max():.4g} mu_s {np.nanmin(out['mu_s']):.4g}..{np.nanmax(out['mu_s']):.4g} "
              f"NaN mu_s {np.isnan(out['mu_s']).mean():.2f} free {np.median(out['free_dist']):.4g} T {out['transmittance'].mean():.4f} "
              f"f {out['f'].mean():.4g} pdf {out['pdf'].mean():.4g} s_pdf {out['s_pdf'].mean():.4g} modes {set(out['s_mode'].tolist())} {set(out['f_mode'].tolist())}")
    np.savez_compressed(os.path.join(HERE, "medium.npz"), **pack)
    print("wrote medium.npz", os.path.getsize(os.path.join(HERE, "medium.npz")) // 1024, "KiB")
