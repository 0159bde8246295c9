"""Known-answer vectors from the reference's own path construction (build container only: needs oracle/_ref/libref_path_halton.so).

    python tests/golden/make_golden_paths.py            -> tests/golden/paths.npz

For the fixture scenes `c10` (static camera) and `motion` (camera + object motion blur) at their fixture frame sizes, Halton points,
--frame 1: oracle/ref_path.c calls the UNMODIFIED reference's path_init + path_extend per path index -- lambda and time sampling,
camera_sample of the thin lens, view_cam_init_frame's slerp, path_propagate -> accel_intersect -- and returns pixel, wavelength, time,
the point on the lens, the ray direction, the camera throughput, the first hit (prim, u, v, dist).  Plus prims_offset_ray on random
points / directions.  One process per scene: the reference keeps its state in the global rt.
"""
import ctypes as C
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REFDIR = os.path.join(ROOT, "oracle", "_ref")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

CASES = ["c10", "motion"]
NEE_CASES = ["c10", "motion", "glass_metal", "sphere_light", "sky_light", "envmap", "fog", "subsurf", "skin"]   # next-event samples at the first hit vertex (ref_path_nee)
N_NEE = 6000
BOUNCE_CASES = ["c10", "glass_metal", "motion", "sphere_light", "fog", "subsurf", "skin", "vstack"]   # the second path_extend (ref_path_bounce), pt and ptdl bookkeeping
N_BOUNCE = 4000
EMISSION_CASES = ["c10", "glass_metal", "motion", "sphere_light", "sky_light", "envmap", "sky_const"]   # what the samplers splat for emission found by extension (ref_path_emission)
N_EMISSION = 20000
SCRAMBLING = 0.5
N_LOW = 6000
SPECIAL = [2**24 - 8, 2**31 - 8, 2**32 - 8, 2**32 + 5, 2**33 + 12345, 123456789012]   # 32-bit clipping of the Halton index, wide indices


def indices():
    return np.concatenate([np.arange(N_LOW, dtype=np.uint64)] + [np.arange(s, s + 16, dtype=np.uint64) for s in SPECIAL])


def worker(case, out):
    from helpers import GoldenImage
    g = GoldenImage(case.split(":")[-1])
    tmp = tempfile.mkdtemp(prefix="corona_paths_")
    nra2 = g.write_files(tmp)
    os.chdir(REFDIR)
    L = C.CDLL(os.path.join(REFDIR, "libref_path_halton.so"), mode=C.RTLD_GLOBAL)   # the shader modules it dlopens resolve rt.* against it
    args = ["-w", str(g.w), "-h", str(g.h), "--frame", "1", "-t", "1", "-s", "1", "-b", "0", "-x"]
    argv = (C.c_char_p * len(args))(*[a.encode() for a in args])
    L.ref_path_open.argtypes = [C.c_char_p, C.c_int, C.POINTER(C.c_char_p)]
    assert L.ref_path_open(nra2.encode(), len(args), argv) == 0
    L.ref_path_camera.argtypes = [C.c_uint64, C.c_uint64, C.c_void_p]
    if case.startswith("nee:"):
        nee = np.zeros((N_NEE, 20), np.float32)
        L.ref_path_nee.argtypes = [C.c_uint64, C.c_uint64, C.c_void_p]
        L.ref_path_nee(0, N_NEE, nee.ctypes.data)
        np.savez(out, nee=nee)
        sys.stdout.flush()
        os._exit(0)
    if case.startswith("bounce:"):
        L.ref_path_bounce.argtypes = [C.c_uint64, C.c_uint64, C.c_float, C.c_int, C.c_void_p]
        res = {}
        for name, with_nee in (("pt", 0), ("ptdl", 1)):
            b = np.zeros((N_BOUNCE, 20), np.float32)
            L.ref_path_bounce(0, N_BOUNCE, SCRAMBLING, with_nee, b.ctypes.data)
            res[name] = b
        np.savez(out, **res)
        sys.stdout.flush()
        os._exit(0)
    if case.startswith("emission:"):
        L.ref_path_emission.argtypes = [C.c_uint64, C.c_uint64, C.c_float, C.c_int, C.c_void_p]
        res = {}
        for name, ptdl in (("pt", 0), ("ptdl", 1)):
            b = np.zeros((N_EMISSION, 8), np.float32)
            L.ref_path_emission(0, N_EMISSION, SCRAMBLING, ptdl, b.ctypes.data)
            keep = (b[:, 3] > 0) | (b[:, 4] > 0)          # only the paths that found an emitter
            res[name] = b[keep]
        np.savez(out, **res)
        sys.stdout.flush()
        os._exit(0)
    idx = indices()
    rows = np.zeros((len(idx), 20), np.float32)
    # consecutive runs of indices in one call each
    start = 0
    while start < len(idx):
        end = start + 1
        while end < len(idx) and idx[end] == idx[end - 1] + 1:
            end += 1
        L.ref_path_camera(int(idx[start]), end - start, rows[start:end].ctypes.data)
        start = end
    rng = np.random.default_rng(5)
    x = (rng.standard_normal((4000, 3)) * np.float32(10.0) ** rng.integers(-3, 4, (4000, 1))).astype(np.float32)
    d = rng.standard_normal((4000, 3)).astype(np.float32)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    off = np.zeros((4000, 4), np.float32)
    L.ref_path_offset.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]
    L.ref_path_offset(x.ctypes.data, d.ctypes.data, 4000, off.ctypes.data)
    np.savez(out, index=idx, rows=rows, off_x=x, off_dir=d, off_out=off)
    sys.stdout.flush()
    os._exit(0)       # the reference's worker pool and display stay up: leave without its atexit handlers


if __name__ == "__main__":
    if len(sys.argv) == 3:
        worker(sys.argv[1], sys.argv[2])
    out = {}
    for case in CASES:
        tmp = tempfile.mktemp(suffix=".npz")
        env = dict(os.environ, LD_LIBRARY_PATH=os.path.join(REFDIR, "shaders") + ":" + os.environ.get("LD_LIBRARY_PATH", ""))   # shader_init dlopens lib<name>.so
        subprocess.run([sys.executable, os.path.abspath(__file__), case, tmp], check=True, stdout=subprocess.DEVNULL, env=env)
        z = np.load(tmp)
        for k in z.files:
            out[f"{case}_{k}"] = z[k]
        os.remove(tmp)
        r = z["rows"]
        print(case, "paths", len(r), "hit fraction", float((r[:, 17] == 2).mean()), "mean lambda", float(r[:, 2].mean()), "time range", float(r[:, 3].min()), float(r[:, 3].max()))
    for case in NEE_CASES:
        tmp = tempfile.mktemp(suffix=".npz")
        env = dict(os.environ, LD_LIBRARY_PATH=os.path.join(REFDIR, "shaders") + ":" + os.environ.get("LD_LIBRARY_PATH", ""))
        subprocess.run([sys.executable, os.path.abspath(__file__), "nee:" + case, tmp], check=True, stdout=subprocess.DEVNULL, env=env)
        nee = np.load(tmp)["nee"]
        os.remove(tmp)
        out[f"{case}_nee"] = nee
        called = nee[:, 4] == 0
        lit = nee[:, 9] > 0
        sky = lit & (nee[:, 10:12].view("u4") == 0xffffffff).all(axis=1)
        print(case, "nee: first hits", int((nee[:, 3] == 2).sum()), "nee_sample called", int(called.sum()), "contributing", int(lit.sum()), "of them sky", int(sky.sum()),
              "mean weight", float(nee[lit, 9].mean()) if lit.any() else 0.0)
    for case in BOUNCE_CASES:
        tmp = tempfile.mktemp(suffix=".npz")
        env = dict(os.environ, LD_LIBRARY_PATH=os.path.join(REFDIR, "shaders") + ":" + os.environ.get("LD_LIBRARY_PATH", ""))
        subprocess.run([sys.executable, os.path.abspath(__file__), "bounce:" + case, tmp], check=True, stdout=subprocess.DEVNULL, env=env)
        z = np.load(tmp)
        for name in ("pt", "ptdl"):
            b = z[name]
            out[f"{case}_bounce_{name}"] = b
            print(case, name, "bounce: first hits", int((b[:, 3] == 2).sum()), "second extend called", int((b[:, 4] != -1).sum()), "went on", int((b[:, 5] == 3).sum()),
                  "second vertex on geometry", int(((b[:, 5] == 3) & ~(b[:, 10:12].view("u4") == 0xffffffff).all(axis=1)).sum()))
        os.remove(tmp)
    for case in EMISSION_CASES:
        tmp = tempfile.mktemp(suffix=".npz")
        env = dict(os.environ, LD_LIBRARY_PATH=os.path.join(REFDIR, "shaders") + ":" + os.environ.get("LD_LIBRARY_PATH", ""))
        subprocess.run([sys.executable, os.path.abspath(__file__), "emission:" + case, tmp], check=True, stdout=subprocess.DEVNULL, env=env)
        z = np.load(tmp)
        for name in ("pt", "ptdl"):
            b = z[name]
            out[f"{case}_emission_{name}"] = b
            print(case, name, "emission by extension: at the first vertex", int((b[:, 3] > 0).sum()), "at the second", int((b[:, 4] > 0).sum()),
                  "mean mis weight there", float(b[b[:, 4] > 0, 5].mean()) if (b[:, 4] > 0).any() else 0.0)
        os.remove(tmp)
    np.savez_compressed(os.path.join(HERE, "paths.npz"), **out)
    print("wrote paths.npz", os.path.getsize(os.path.join(HERE, "paths.npz")) // 1024, "KiB")
