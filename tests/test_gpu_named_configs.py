"""BASELINE.json's named configurations AT THEIR STATED SIZES, judged with the reference's own regression criterion
(run with -m gpu on the B200 box).

    regression/0010_pt     pt,   1024x576, 128 spp, maxerror 4.0     -> conv_c10_pt.npz
    regression/0011_ptdl   ptdl, 1024x576, 128 spp, maxerror 3.8     -> conv_c10_ptdl.npz
    regression/0002_mb     ptdl + motion blur (Halton, rec709), 512x288, 128 spp, maxerror 0.11 -> conv_motion.npz

The reference's test is `pfmdiff testrender.pfm reference.pfm <= maxerror` (regression/createres.sh:20-30,
tools/img/pfmdiff.c:75-86: sqrt of the summed squared channel differences per pixel), reference.pfm being a long render that
is not available offline.  tests/golden/make_golden_converged.py made those long renders with the UNMODIFIED reference renderer
(oracle/_ref, 4096 / 8192 spp, another --frame) together with the RMSE of the reference's own 32 / 128 / 512-spp renders
against them.  Here the GPU renders the same scenes at the same sizes through the C ABI and must
  (1) pass the reference's criterion at the regression's sample count (128 spp),
  (2) land within 15 % of the RMSE the reference renderer itself reaches at 32, 128 and 512 spp, and
  (3) converge like it: rand variants fall as 1/sqrt(spp) (a factor 2 per 4x samples, within 15 %; the long render's own
      residual noise bends that slightly, for the reference too), the Halton variant at least that fast.
"""
import os

import numpy as np
import pytest

from helpers import GoldenImage

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = ["c10_pt", "c10_ptdl", "motion"]
SPP = (32, 128, 512)


def pfmdiff_rmse(a, b):
    """tools/img/pfmdiff.c:75-86"""
    d = a.astype(np.float64) - b.astype(np.float64)
    return float(np.sqrt((d * d).sum() / (a.shape[0] * a.shape[1])))


@pytest.fixture(scope="module")
def gpu(lib):
    if lib.device_count() < 1:
        pytest.fail("no CUDA device: " + lib.load().cb200_last_error().decode())
    lib.set_device(0)
    return lib


@pytest.mark.parametrize("case", CASES)
def test_named_config_at_stated_size(gpu, case):
    z = np.load(os.path.join(GOLDEN, f"conv_{case}.npz"))
    g = GoldenImage(str(z["fixture"]))
    w, h, key = int(z["w"]), int(z["h"]), str(z["key"])
    long = z["long"].astype(np.float32)
    assert int(z["clipped"]) == 0 and list(z["ref_spp"]) == list(SPP)
    ref_rmse = z["ref_rmse"].mean(axis=1)
    acc = gpu.Accel(g.scene).build()
    r = gpu.Render(acc, g.camera, g.materials, w, h, frame=1, **g.sky_args, **GoldenImage.variant_args(key))
    got = {}
    for spp in range(1, SPP[-1] + 1):
        r.render_pass()
        if spp in SPP:
            img = r.image()
            assert img.shape == long.shape and np.isfinite(img).all()
            got[spp] = pfmdiff_rmse(img, long)
    means = r.image().astype(np.float64).mean(axis=(0, 1)) / long.astype(np.float64).mean(axis=(0, 1))
    st = r.stats()
    r.close()
    acc.close()
    print(f"{case} ({key}, {w}x{h}): pfmdiff RMSE vs the {int(z['long_spp'])}-spp reference render: "
          + ", ".join(f"{s} spp {got[s]:.4f} (reference renderer {ref_rmse[i]:.4f})" for i, s in enumerate(SPP))
          + f"; maxerror {float(z['maxerror'])}; channel means / long {np.round(means, 4)}")
    assert st["paths"] == SPP[-1] * w * h
    # (1) the regression's own pass criterion at its own sample count
    assert got[128] <= float(z["maxerror"]), f"{case}: RMSE {got[128]:.4f} at 128 spp exceeds regression/*/maxerror {float(z['maxerror'])}"
    # (2) the same distance from the converged image as the reference renderer at every sample count
    for i, s in enumerate(SPP):
        assert abs(got[s] / ref_rmse[i] - 1) < 0.15, f"{case}: RMSE {got[s]:.4f} at {s} spp, the reference renderer reaches {ref_rmse[i]:.4f}"
    # (3) convergence rate
    for a, b in ((32, 128), (128, 512)):
        ratio = got[a] / got[b]
        if "halton" in key:
            assert ratio > 2.0 * 0.85, f"{case}: RMSE falls by {ratio:.2f} from {a} to {b} spp"
        else:
            assert abs(ratio / 2.0 - 1) < 0.15, f"{case}: RMSE falls by {ratio:.2f} from {a} to {b} spp, 1/sqrt(spp) predicts 2"
    assert np.all(np.abs(means - 1) < 0.01), f"{case}: channel means off the converged image: {means}"
