"""scratch: render one golden case/variant on the GPU and save the image for offline comparison"""
import sys, os, importlib
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
from helpers import GoldenImage
lib = importlib.import_module("corona-13_b200.lib")
case, key = sys.argv[1], sys.argv[2]
g = GoldenImage(case)
acc = lib.Accel(g.scene).build()
img, st = g.render(lib, acc, key)
np.save(os.path.join(ROOT, "gpurun_out", f"img_{case}_{key}.npy"), img)
print(st)
