"""Material table of bench.py's synthetic 10 M-triangle config (terrain + icosphere soup + quad light):
the shader list below flattened through the reference's rgb -> spectrum coefficient table (which only exists where
oracle/_ref was built), stored so that bench.py runs anywhere.   python scripts/make_bench_materials.py"""
import importlib
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
IO = importlib.import_module("corona-13_b200").scene_io

LINES = ["diffuse",                  # 0
         "color d 0.55 0.5 0.4",     # 1
         "mult 1 1 0",               # 2 terrain
         "color d 0 0 0",            # 3
         "color e 60 60 60 1.",      # 4
         "mult 2 3 4 0",             # 5 light
         "color d 0.3 0.45 0.7",     # 6
         "mult 1 6 0"]               # 7 soup
SHAPE_MATS = [2, 7, 5]               # scenes.synthetic_scene: terrain, soup, light

if __name__ == "__main__":
    tmp = tempfile.mkdtemp()
    nra2 = os.path.join(tmp, "bench.nra2")
    IO.write_nra2(nra2, LINES, [(m, f"shape{i}") for i, m in enumerate(SHAPE_MATS)])
    ms, _, _ = IO.parse_nra2(nra2, IO.Rgb2Spec(IO.coeff_path(ROOT)))
    mats, _ = ms.carrays()
    np.savez_compressed(os.path.join(ROOT, "corona-13_b200", "data", "bench_materials.npz"), materials=np.frombuffer(bytes(mats), np.uint8),
                        shader_lines=np.array(LINES), shape_mats=np.int64(SHAPE_MATS))
    print("wrote bench_materials.npz", len(ms.materials), "materials")
