"""CPU-side checks of the reference-render fixtures themselves (no GPU): the properties the GPU tests lean on must already hold
for the reference's own images, otherwise a GPU test that passes proves nothing.

  * dbor.npz  : the cascade levels of view_splat_col (src/view.c:497-522) split every sample between two levels with weights
                lv + uv = 1, so the reference's levels must sum to its framebuffer; levels fill from the dim end;
  * ptnee.npz : ptnee (src/sampler.d/ptnee.c) drops emission found by extension, so away from directly visible emitters its
                images equal ptdl's in expectation and are never brighter overall;
  * img_sphere_light.npz : upstream's sphere-light sampling covers the upper hemisphere only (prims.c:228), which makes its ptdl
                darker than its pt on this scene -- the fixture must show that bias, it is what the GPU reproduces."""
import os

import numpy as np

from helpers import GoldenImage

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_dbor_levels_sum_to_the_framebuffer():
    z = np.load(os.path.join(GOLDEN, "dbor.npz"))
    for run in [str(x) for x in z["runs"]]:
        case, key, levels = run.split(":")
        for seed in (1, 2):
            fb, lv = z[f"{case}_{key}_fb_seed{seed}"].astype(np.float64), z[f"{case}_{key}_dbor_seed{seed}"].astype(np.float64)
            assert lv.shape == (int(levels), ) + fb.shape
            assert np.all(lv >= 0)
            assert np.all(np.abs(lv.sum(axis=0) - fb) <= 2e-2 * np.maximum(fb, fb.mean())), run
            means = lv.mean(axis=(1, 2, 3))
            filled = np.nonzero(means > 0)[0]
            assert filled[0] == 0 and np.array_equal(filled, np.arange(len(filled))), f"{run}: levels fill from the dim end: {means}"
        # the framebuffer of the --dbor run is the plain image fixture's render (same scene, same Halton points)
        a = GoldenImage(case).ref(key, 1).astype(np.float64)
        fb = z[f"{case}_{key}_fb_seed1"].astype(np.float64)
        assert abs(fb.mean() / a.mean() - 1) < 1e-3


def test_ptnee_is_ptdl_minus_the_extension_half():
    z = np.load(os.path.join(GOLDEN, "ptnee.npz"))
    for case in [str(x) for x in z["cases"]]:
        g = GoldenImage(case)
        nee, ptdl = z[f"{case}_seed1"].astype(np.float64), g.ref("ptdl_halton", 1).astype(np.float64)
        assert nee.shape == ptdl.shape and np.isfinite(nee).all()
        assert nee.mean() <= 1.01 * ptdl.mean(), case
        assert nee.mean() >= 0.7 * ptdl.mean(), case


def test_sphere_light_fixture_shows_the_upstream_sampling_bias():
    g = GoldenImage("sphere_light")
    pt = 0.5 * (g.ref("pt_halton", 1).astype(np.float64).mean() + g.ref("pt_halton", 2).astype(np.float64).mean())
    ptdl = 0.5 * (g.ref("ptdl_halton", 1).astype(np.float64).mean() + g.ref("ptdl_halton", 2).astype(np.float64).mean())
    assert 0.80 < ptdl / pt < 0.95, ptdl / pt
