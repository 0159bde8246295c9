"""The plain-C scene readers under AddressSanitizer + UndefinedBehaviorSanitizer (CPU only): the host sources are compiled once with
-fsanitize=address,undefined and fed the fixture scenes plus seeded random corruptions of a .geo file (truncations, header bytes,
payload bytes, wild 32-bit words).  Parsing stops before any GPU work (--dump-materials), so this runs without a device.  Every
run must end with exit code 0 (loaded, possibly with a shape dropped) or 2 (refused) and without a sanitizer report."""
import os
import random
import struct
import subprocess

import pytest

from helpers import GoldenImage

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "corona-13_b200")
COEFF = os.path.join(ROOT, "oracle", "_ref", "data", "ergb2spec.coeff")
TABLES = os.path.join(ROOT, "corona-13_b200", "data", "ref_tables.cbt")

pytestmark = pytest.mark.skipif(not os.path.exists(COEFF), reason="data/ergb2spec.coeff only exists where oracle/_ref was built")


@pytest.fixture(scope="module")
def asan_cli(built, tmp_path_factory):
    out = str(tmp_path_factory.mktemp("asan") / "corona_b200_asan")
    src = [os.path.join(PKG, "host", f) for f in ("main_b200.c", "scene_b200.c", "accel_b200.c", "render_b200.c")]
    cmd = ["gcc", "-std=c11", "-D_GNU_SOURCE", "-O1", "-g", "-fsanitize=address,undefined", "-fno-omit-frame-pointer",
           "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(PKG, "host"), *src, "-o", out, "-L" + PKG, "-lcorona_b200", "-Wl,-rpath," + PKG, "-lm"]
    p = subprocess.run(cmd, capture_output=True, text=True)
    if p.returncode != 0:
        pytest.skip("no sanitizer runtime for gcc here: " + p.stderr[-200:])
    return out


def run(binary, nra2, tmp):
    env = dict(os.environ, ASAN_OPTIONS="detect_leaks=1", UBSAN_OPTIONS="print_stacktrace=1")
    p = subprocess.run([binary, nra2, "--coeff", COEFF, "--tables", TABLES, "--dump-materials", os.path.join(tmp, "m.bin")], capture_output=True, text=True, env=env)
    clean = "Sanitizer" not in p.stderr and "runtime error" not in p.stderr
    return p.returncode, clean, p.stderr[-600:]


@pytest.mark.parametrize("case", ["c10", "glass_metal", "skin", "envmap", "vstack", "motion"])
def test_fixture_scenes_parse_clean_under_sanitizers(asan_cli, tmp_path, case):
    nra2 = GoldenImage(case).write_files(str(tmp_path))
    rc, clean, err = run(asan_cli, nra2, str(tmp_path))
    assert rc == 0 and clean, err


def test_corrupted_geo_files_under_sanitizers(asan_cli, tmp_path):
    g = GoldenImage("motion")      # triangles, quads-free soup with motion blur, an analytic sphere, a quad light
    nra2 = g.write_files(str(tmp_path))
    rnd = random.Random(7)
    for shape in ("shape0", "shape2"):
        geo = str(tmp_path / (shape + ".geo"))
        orig = open(geo, "rb").read()
        for trial in range(24):
            h = bytearray(orig)
            kind = trial % 4
            if kind == 0:
                h = h[:rnd.randrange(0, len(h))]
            elif kind == 1:
                for _ in range(rnd.randrange(1, 6)):
                    h[rnd.randrange(0, min(len(h), 64))] = rnd.randrange(256)
            elif kind == 2:
                for _ in range(rnd.randrange(1, 20)):
                    h[rnd.randrange(0, len(h))] = rnd.randrange(256)
            else:
                off = rnd.randrange(0, len(h) - 8)
                h[off:off + 4] = struct.pack("<I", rnd.choice([0xffffffff, 0x7fffffff, 0x80000000, 1 << 28]))
            open(geo, "wb").write(bytes(h))
            rc, clean, err = run(asan_cli, nra2, str(tmp_path))
            assert rc in (0, 2) and clean, f"{shape} trial {trial} (kind {kind}): rc {rc}\n{err}"
        open(geo, "wb").write(orig)


def test_corrupted_scene_lists_under_sanitizers(asan_cli, tmp_path):
    """.nra2-level mutations: wild shape counts (negative, 4 * 10^9), wild shader counts, `mult` chains that fan out
    exponentially, forward and self references, BSDF shaders in pre slots, lines cut short"""
    g = GoldenImage("glass_metal")
    nra2 = g.write_files(str(tmp_path))
    lines = open(nra2).read().split("\n")
    nsh = int(lines[1])
    shape_line = 2 + nsh

    def variant(name, mutate):
        l = list(lines)
        mutate(l)
        path = str(tmp_path / f"{name}.nra2")
        open(path, "w").write("\n".join(l))
        return path

    cases = {
        "shapes_negative": lambda l: l.__setitem__(shape_line, "-1"),
        "shapes_huge": lambda l: l.__setitem__(shape_line, "4000000000"),
        "shapes_more_than_listed": lambda l: l.__setitem__(shape_line, str(int(l[shape_line]) + 1000)),
        "shaders_negative": lambda l: l.__setitem__(1, "-5"),
        "shaders_huge": lambda l: l.__setitem__(1, "100000"),
        "mult_self": lambda l: l.__setitem__(2 + nsh - 1, "mult 1 0 0"),
        "mult_forward": lambda l: l.__setitem__(2, f"mult 1 {nsh - 1} {nsh - 1}"),
        "truncated": lambda l: l.__delitem__(slice(4, None)),
    }
    for name, mutate in cases.items():
        import time
        t0 = time.time()
        rc, clean, err = run(asan_cli, variant(name, mutate), str(tmp_path))
        assert rc in (0, 2) and clean, f"{name}: rc {rc}\n{err}"
        assert time.time() - t0 < 20, f"{name}: the reader took {time.time() - t0:.1f} s"
    # 17 lines of `mult 16 -1 ... -1`: 16^16 prepare steps if followed naively
    fan = ["black", "18", "diffuse"] + ["mult 16 " + " ".join(["-1"] * 17)] * 17 + ["1", "17 shape0"]
    path = str(tmp_path / "fanout.nra2")
    open(path, "w").write("\n".join(fan) + "\n")
    import time
    t0 = time.time()
    rc, clean, err = run(asan_cli, path, str(tmp_path))
    assert rc in (0, 2) and clean and time.time() - t0 < 20, f"fan-out: rc {rc} after {time.time() - t0:.1f} s\n{err}"
