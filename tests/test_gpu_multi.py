"""Multi-GPU through the plain-C host (run with -m gpu on a box with >= 2 GPUs: `gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu`;
skipped on a single GPU).  `corona_b200 --gpus N` forks one process per GPU, splits the progressions between them and sums the
framebuffers into rank 0 with one NCCL reduce per progression group (csrc/comm.cu).  Every random draw is a function of
(frame, path index, dimension), so the N-GPU image must equal the 1-GPU image of the same command line up to fp32 summation order."""
import os
import subprocess

import numpy as np
import pytest

from helpers import GoldenImage, cb

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "corona-13_b200", "corona_b200")
COEFF = os.path.join(ROOT, "oracle", "_ref", "data", "ergb2spec.coeff")
TABLES = os.path.join(ROOT, "corona-13_b200", "data", "ref_tables.cbt")


def render(nra2, out, *args):
    p = subprocess.run([BIN, nra2, "--coeff", COEFF, "--tables", TABLES, "-x", out, *args], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr[-2000:] + p.stdout[-1000:]
    return cb.scene_io.read_pfm(os.path.splitext(nra2)[0] + out + "_fb00.pfm")


@pytest.mark.parametrize("case,key,batch", [("c10", "ptdl_halton", "1"), ("glass_metal", "ptdl_rand", "3"), ("fog", "ptdl_halton", "1")])
def test_n_gpu_image_equals_one_gpu_image(built, lib, tmp_path, case, key, batch):
    n = lib.device_count()
    if n < 2:
        pytest.skip("needs at least two GPUs")
    g = GoldenImage(case)
    nra2 = g.write_files(str(tmp_path))
    f = key.split("_")
    common = ("-s", "37", "-w", str(g.w), "-h", str(g.h), "--frame", "1", "--sampler", f[0], "--points", f[1], "--batch", batch)
    one = render(nra2, "one", *common)
    for gpus in sorted({2, min(n, 8)}):
        many = render(nra2, f"n{gpus}", *common, "--gpus", str(gpus))
        scale = float(one.astype(np.float64).mean())
        d = (many.astype(np.float64) - one.astype(np.float64))
        rel_rmse = float(np.sqrt((d ** 2).mean())) / scale
        assert rel_rmse < 1e-5, f"{case}/{key} on {gpus} GPUs: relative RMSE {rel_rmse:.3e} against the 1-GPU image"
        assert float(np.abs(d).max()) / max(float(one.max()), 1e-30) < 1e-4
