"""Known-answer vectors for the BSDFs (BASELINE configs[3] = regression/0052_dielectric, 0053): the
reference's OWN sample()/brdf()/pdf() callbacks -- shaders/libdielectric.so, libmetal.so and the built-in diffuse of
src/shader.c -- driven by oracle/ref_bsdf.c with the path set up like tools/battle-test.c:57-140.  Build container only:

    python tests/golden/make_golden_bsdf.py      ->  tests/golden/bsdf.npz

Cases: the battle-test configuration itself (lambda 525 nm, roughness 0.4, "dielectric 1.7 73", incident angles u = k/3.5,
both the reflect and the transmit set-up) plus random incident / outgoing directions, wavelengths and random numbers for
rough + smooth dielectric, rough + polished metal (Au, Ag) and diffuse.
"""
import ctypes as C
import importlib
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
IO = importlib.import_module("corona-13_b200").scene_io
REFDIR = os.path.join(ROOT, "oracle", "_ref")

# name -> (shader module, rest of its .nra2 line, roughness values to exercise)
CASES = {"dielectric": ("libdielectric.so", "1.7 73", [0.4, 0.05, 0.0]),
         "dielectric_c10": ("libdielectric.so", "1.3 23", [0.04]),
         "metal_au": ("libmetal.so", "Au", [0.2, 0.0]),
         "metal_ag": ("libmetal.so", "Ag", [0.1]),
         "diffuse": ("diffuse", "", [1.0]),
         "diffdiel": ("libdiffdiel.so", "1.33 30", [0.13, 0.4, 0.0])}     # regression/0030_subsurf: `diffdiel 1.33 30` under roughness 0.13


def unit(v):
    return (v / np.linalg.norm(v, axis=1)[:, None]).astype(np.float32)


def queries(rng, roughness, n_random=1500):
    q = []
    for flip in (0, 1):                                  # battle-test.c:108-116, REFLECT and !REFLECT
        for k in range(4):
            u = np.float32(k / 3.5)
            wi = np.float32([0.0, np.sqrt(u), (1.0 if flip else -1.0) * np.sqrt(1 - u)])
            for _ in range(200):
                wo = rng.normal(size=3)
                wo[2] = abs(wo[2])                           # battle-test only looks at the +z hemisphere
                q.append((wi, wo / np.linalg.norm(wo), 525.0, rng.random(3), flip))
    for _ in range(n_random):
        flip = int(rng.integers(0, 2))
        wi = rng.normal(size=3)
        wi[2] = (1 if flip else -1) * abs(wi[2])            # arriving from the side the normal points to
        wo = rng.normal(size=3)
        q.append((wi / np.linalg.norm(wi), wo / np.linalg.norm(wo), rng.uniform(380.0, 780.0), rng.random(3), flip))
    out = np.zeros(len(q), IO.BSDF_QUERY)
    for i, (wi, wo, lam, r, flip) in enumerate(q):
        out[i]["wi"], out[i]["wo"], out[i]["lambda_"], out[i]["rand"], out[i]["flip"] = wi, wo, lam, r, flip
    out["rd"], out["rs"], out["rg"], out["roughness"] = 0.8, 0.06, 1.0, roughness     # battle-test.c:98-103
    return out


def battle_queries(flip, k, size=512, spp=8):
    """tools/battle-test.c:108-211 for incident angle k: spp*size^2 sample() draws + one brdf()/pdf() evaluation per disk pixel"""
    n = spp * size * size
    rng = np.random.default_rng(666 + flip)
    for _ in range(k):
        rng.random((n, 3), dtype=np.float32)                  # angle k uses the k-th block of the stream
    j, i = np.meshgrid(np.arange(size), np.arange(size), indexing="ij")
    x = (2.0 * i / np.float32(size) - 1.0).astype(np.float32).reshape(-1)
    y = (2.0 * j / np.float32(size) - 1.0).astype(np.float32).reshape(-1)
    len2 = x * x + y * y
    u = np.float32(k / 3.5)
    q = np.zeros(n, IO.BSDF_QUERY)
    q["wi"] = np.float32([0.0, np.sqrt(u), (1.0 if flip else -1.0) * np.sqrt(1 - u)])
    q["lambda_"], q["rd"], q["rs"], q["rg"], q["roughness"], q["flip"] = 525.0, 0.8, 0.06, 1.0, 0.4, flip
    q["rand"] = rng.random((n, 3), dtype=np.float32)
    q["wo"] = np.tile(np.stack([x, y, np.sqrt(np.maximum(0.0, 1.0 - len2))], -1).astype(np.float32), (spp, 1))
    return q, len2 < 1.0


def battle_sums(out, inside, size=512):
    """(ebsdf, bsdf, epdf, pdf) as battle-test prints them (tools/battle-test.c:155-162,213-219)"""
    n = len(out)
    ok = (out["s_wo"][:, 2] > 0) & (out["s_weight"] > 0)
    g = slice(0, size * size)
    return (float(out["s_weight"][ok].astype(np.float64).sum() / n),
            float((out["f"][g][inside].astype(np.float64) * 4.0 / (size * size)).sum()),
            float(ok.sum() / n),
            float((out["pdf"][g][inside].astype(np.float64) * 4.0 / (size * size)).sum()))


if __name__ == "__main__":
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "ref"])
    L = C.CDLL(os.path.join(REFDIR, "libref_bsdf.so"), mode=C.RTLD_GLOBAL)   # the shader modules resolve pointsampler / rt / path_eta_ratio against it
    L.ref_bsdf_open.argtypes = [C.c_char_p, C.c_char_p]
    L.ref_bsdf_eval.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_uint64]
    rng = np.random.default_rng(52)
    pack = {"cases": np.array(list(CASES))}
    for name, (so, line, roughs) in CASES.items():
        path = so if so == "diffuse" else os.path.join(REFDIR, "shaders", so)
        h = L.ref_bsdf_open(path.encode(), line.encode())
        assert h >= 0, name
        q = np.concatenate([queries(rng, r) for r in roughs])
        out = np.zeros(len(q), IO.BSDF_RESULT)
        L.ref_bsdf_eval(h, q.ctypes.data, out.ctypes.data, len(q))
        pack[name + "_q"], pack[name + "_r"] = q.view("u1").reshape(len(q), -1), out.view("u1").reshape(len(q), -1)
        pack[name + "_line"] = np.array(line)
        ok = out["s_weight"] > 0
        print(f"{name}: {len(q)} queries, sample() > 0 in {ok.mean():.2f}, brdf() > 0 in {(out['f'] > 0).mean():.2f}, pdf() > 0 in {(out['pdf'] > 0).mean():.2f}, "
              f"mean weight {out['s_weight'][ok].mean():.4f}, NaNs {np.isnan(out['s_weight']).sum() + np.isnan(out['f']).sum()}")
    # regression/0052_dielectric (reflect) and 0053 (transmit) run on the reference's own module
    h = L.ref_bsdf_open(os.path.join(REFDIR, "shaders", "libdielectric.so").encode(), b"1.7 73")
    sums = np.zeros((2, 4, 4))
    for flip in (0, 1):
        for k in range(4):
            q, inside = battle_queries(flip, k)
            out = np.zeros(len(q), IO.BSDF_RESULT)
            L.ref_bsdf_eval(h, q.ctypes.data, out.ctypes.data, len(q))
            sums[flip, k] = battle_sums(out, inside)
            print("battle-test reference", "transmit" if flip else "reflect", k, sums[flip, k], flush=True)
    pack["battle_sums"] = sums
    np.savez_compressed(os.path.join(HERE, "bsdf.npz"), **pack)
    print("wrote bsdf.npz", os.path.getsize(os.path.join(HERE, "bsdf.npz")) // 1024, "KiB")
