"""Measured lookup tables the reference's shader modules carry as C initialisers -- extracted (build container only) into
tests/golden/ref_tables.npz so tests and bench can hand them to cb200_render_create as cb_table_t:

  checker   140 x 36  ColorChecker SG reflectances, 380 nm + 10 nm steps   src/shaders/colorcheckersg.c:48-193
  metal_<m>   2 x 95  n and k of Ti Cu Fe Au Ag, 360 nm + 5 nm steps        src/shaders/fresnel.h:21-516

    python tests/golden/make_tables.py
"""
import os
import re

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("CORONA_REF", "/root/reference")
NUM = r"[-+]?(?:\d+\.\d*|\.\d+|\d+)(?:[eE][-+]?\d+)?"


def floats(text):
    return np.array([float(x) for x in re.findall(NUM, re.sub(r"(?<=[\d.])f", "", text))], np.float32)


def main():
    src = open(os.path.join(REF, "src/shaders/colorcheckersg.c")).read()
    body = src[src.index("static const float cobs[140][36]"):src.index("// END_DATA")]
    body = body[body.index("{"):]
    checker = floats(body).reshape(140, 36)
    src = open(os.path.join(REF, "src/shaders/fresnel.h")).read()
    names = re.findall(r'"(\w+)"', src[src.index("fresnel_ior_material[]"):src.index("};", src.index("fresnel_ior_material[]"))])
    body = src[src.index("static const float fresnel_ior[][95][2]"):src.index("static inline void fresnel_get_ior_mf")]
    body = re.sub(r"//[^\n]*", "", body[body.index("{"):])
    nk = floats(body).reshape(len(names), 95, 2)
    out = {"checker": checker}
    for i, n in enumerate(names):
        out["metal_" + n.lower()] = np.ascontiguousarray(nk[i].T)   # rows: n, k
    path = os.path.join(HERE, "ref_tables.npz")
    np.savez_compressed(path, **out)
    # the same tables for the C host layer (host/scene_b200.c: tables_load): "CBT1", count, {name[32], rows, cols, lambda_min, lambda_step, data}
    import struct
    cbt = os.path.join(os.path.dirname(os.path.dirname(HERE)), "corona-13_b200", "data", "ref_tables.cbt")   # product data: beside the library
    with open(cbt, "wb") as f:
        f.write(b"CBT1" + struct.pack("<I", len(out)))
        for k, v in out.items():
            lmin, step = (380.0, 10.0) if k == "checker" else (360.0, 5.0)
            f.write(k.encode().ljust(32, b"\0") + struct.pack("<IIff", v.shape[0], v.shape[1], lmin, step) + np.ascontiguousarray(v, "<f4").tobytes())
    print("wrote", path, {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
