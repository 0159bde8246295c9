"""Known-answer vectors of the path construction (run with -m gpu): tests/golden/paths.npz holds what the UNMODIFIED reference's
path_init + path_extend computed per path index (oracle/ref_path.c, tests/golden/make_golden_paths.py) -- pixel, wavelength, time,
the sampled point on the lens, the primary ray direction, the camera throughput, the first hit -- for the fixture scenes `c10`
and `motion` (camera + object motion blur) with Halton points, and prims_offset_ray on random inputs.  The GPU's camera kernel
(k_path_start through cb200_render_camera_rays), its traversal and its ray offset must reproduce them: rows SURVEY 8 a13 / a18."""
import os

import numpy as np
import pytest

from helpers import GoldenImage, R

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def gpu(lib):
    if lib.device_count() < 1:
        pytest.fail("no CUDA device: " + lib.load().cb200_last_error().decode())
    lib.set_device(0)
    return lib


def runs(idx):
    start = 0
    while start < len(idx):
        end = start + 1
        while end < len(idx) and idx[end] == idx[end - 1] + 1:
            end += 1
        yield start, end
        start = end


@pytest.mark.parametrize("case", ["c10", "motion"])
def test_camera_sample_and_first_hit_match_the_reference(gpu, case):
    z = np.load(os.path.join(GOLDEN, "paths.npz"))
    idx, want = z[f"{case}_index"], z[f"{case}_rows"]
    g = GoldenImage(case)
    acc = gpu.Accel(g.scene).build()
    r = gpu.Render(acc, g.camera, g.materials, g.w, g.h, frame=1, **g.sky_args, **GoldenImage.variant_args("pt_halton"))
    rays = np.zeros(len(idx), R.RAY)
    aux = np.zeros((len(idx), 4), np.float32)
    for a, b in runs(idx):
        rays[a:b], aux[a:b] = r.camera_rays(int(idx[a]), b - a)
    # pixel, wavelength, time: pure functions of the Halton points -- bit-exact
    assert np.array_equal(aux[:, 0].view("u4"), want[:, 0].view("u4")) and np.array_equal(aux[:, 1].view("u4"), want[:, 1].view("u4")), "pixel"
    assert np.array_equal(aux[:, 2].view("u4"), want[:, 2].view("u4")), "wavelength"
    assert np.array_equal(rays["time"].view("u4"), want[:, 3].view("u4")), "time"
    # lens point, direction, camera throughput go through sinf/cosf/sqrtf of the lens sample and the slerp of the moving camera (libm ulps)
    assert np.allclose(rays["pos"], want[:, 4:7], rtol=2e-6, atol=2e-6), np.abs(rays["pos"] - want[:, 4:7]).max()
    assert np.allclose(rays["dir"], want[:, 7:10], rtol=0, atol=3e-6), np.abs(rays["dir"] - want[:, 7:10]).max()
    assert np.allclose(aux[:, 3], want[:, 11], rtol=1e-5), np.abs(aux[:, 3] / want[:, 11] - 1).max()
    exact = (rays["pos"].view("u4") == want[:, 4:7].view("u4")).all(axis=1) & (rays["dir"].view("u4") == want[:, 7:10].view("u4")).all(axis=1)
    print(f"{case}: {len(idx)} camera samples, {exact.mean():.3f} of the rays bit-identical to the reference's, the rest within 3e-6")
    # first hit of those rays: the reference's accel_intersect answer (its own tree) against ours
    hits = acc.intersect(rays)
    prim_want = np.ascontiguousarray(want[:, 13:15]).view("u4")
    prim_got = np.ascontiguousarray(hits["prim"]).reshape(-1, 2)
    same = (prim_got == prim_want).all(axis=1)
    assert same.mean() > 0.995, f"{case}: only {same.mean():.4f} of the first hits agree"
    ok = same & exact
    assert np.array_equal(hits["dist"][ok].view("u4"), want[ok, 10].view("u4")), "e[1].dist on bit-identical rays"
    assert np.allclose(hits["dist"][same], want[same, 10], rtol=1e-5)
    vc = R.primid_vcnt(R.hit_prim64(hits))
    tri = ok & ((vc == R.PRIM_TRI) | (vc == R.PRIM_QUAD)) & ~((prim_want[:, 0] == 0xffffffff) & (prim_want[:, 1] == 0xffffffff))   # analytic prims: u,v through libm
    assert np.array_equal(hits["u"][tri].view("u4"), want[tri, 15].view("u4")) and np.array_equal(hits["v"][tri].view("u4"), want[tri, 16].view("u4"))
    # prims_offset_ray
    got = r.offset_ray(z[f"{case}_off_x"], z[f"{case}_off_dir"])
    assert np.array_equal(got.view("u4"), z[f"{case}_off_out"][:, :3].view("u4")), "prims_offset_ray"
    assert (z[f"{case}_off_out"][:, 3] == 0.0).all()
    r.close()
    acc.close()
