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


def test_vertex_level_vectors_are_self_consistent():
    """paths.npz (oracle/ref_path.c through tests/golden/make_golden_paths.py): what the vertex-level GPU tests lean on must hold for
    the reference's own numbers: mis weights in (0, 1], pt's emission weight exactly 1 and the same emitters found by pt and ptdl at
    the first vertex, unit directions, pt and ptdl sampling the SAME second edge wherever both go on (nee_sample + path_pop must only
    move the random dimensions it folds into the vertex, pathspace.c:298 -- the bookkeeping the device reproduces)."""
    z = np.load(os.path.join(GOLDEN, "paths.npz"))
    nee = [k for k in z.files if k.endswith("_nee")]
    assert len(nee) >= 9
    for k in nee:
        a = z[k]
        lit = a[:, 9] > 0
        assert lit.sum() > 1000, k
        assert np.all(a[lit, 9] <= 1.0) and np.all(a[lit, 6] > 0) and np.all(a[lit, 7] > 0), k
        d = a[lit, 15:18].astype(np.float64)
        assert np.allclose((d*d).sum(axis=1), 1.0, atol=1e-5), k
        assert np.all(a[lit, 18] > 0), k                                  # distance to the sampled point
        assert np.all(a[~lit, 10:19] == 0), k                             # nothing recorded for samples that add nothing
    for k in [k for k in z.files if k.endswith("_bounce_pt")]:
        pt, dl = z[k], z[k[:-2] + "ptdl"]
        assert np.array_equal(pt[:, :3].view("u4"), dl[:, :3].view("u4")), k      # same pixels / wavelengths: same camera samples
        both = (pt[:, 5] == 3) & (dl[:, 5] == 3)
        assert both.sum() > 1500, k
        d = pt[both, 6:9].astype(np.float64)
        assert np.allclose((d*d).sum(axis=1), 1.0, atol=1e-5), k
        # the two samplers use different random dimensions for this edge (ptdl's vertex has folded four more in): the directions
        # differ -- checked on the scenes of rough surfaces / media; on polished ones the direction does not depend on the sample
        if k.split("_bounce_")[0] in ("c10", "motion", "sphere_light", "fog"):
            assert (np.abs(pt[both, 6:9] - dl[both, 6:9]).max(axis=1) > 1e-3).mean() > 0.9, k
        assert np.array_equal(pt[both, 12:15].view("u4"), dl[both, 12:15].view("u4")), k   # ... from the same first vertex
    for k in [k for k in z.files if k.endswith("_emission_pt")]:
        pt, dl = z[k], z[k[:-2] + "ptdl"]
        assert np.all(pt[pt[:, 4] > 0, 5] == 1.0), k
        w = dl[dl[:, 4] > 0, 5]
        assert len(w) == 0 or (np.all(w > 0) and np.all(w <= 1.0) and w.mean() < 0.99), k   # ptdl weighs extension against next events
        first = lambda a: {tuple(x) for x in np.ascontiguousarray(a[a[:, 3] > 0][:, :3]).view("u4")}
        assert first(pt) == first(dl), k                                   # the camera sees the same emitters in both
