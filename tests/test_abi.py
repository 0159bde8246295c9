"""The C-ABI library loads without a GPU, exports every symbol include/corona_b200.h declares, and fails
loudly (no CPU fallback) when there is no device."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from helpers import S

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols(header):
    txt = open(os.path.join(ROOT, "include", header)).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(cb200_\w+)\s*\(", txt)))


def test_exports_every_declared_symbol(lib):
    L = lib.load()
    syms = declared_symbols("corona_b200.h")
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(L, s), f"libcorona_b200.so does not export {s}"
    assert sorted(lib.SYMBOLS) == syms
    rsyms = declared_symbols("corona_b200_render.h")
    assert len(rsyms) >= 9
    for s in rsyms:
        assert hasattr(L, s), f"libcorona_b200.so does not export {s}"
    assert sorted(lib.RENDER_SYMBOLS) == rsyms


def test_no_cpu_fallback(lib):
    """without a device every compute entry must refuse; with one this test is a no-op"""
    if lib.device_count() > 0:
        pytest.skip("a GPU is present")
    sc = S.synthetic_scene(200, seed=1)
    with pytest.raises(lib.Cb200Error):
        lib.Accel(sc)
    assert lib.load().cb200_last_error() != b""


def test_host_layer_exports_accel_h(built):
    """libcorona_host.so = the reference's accel.h surface (include/accel.h:28-50) + the batched extensions"""
    L = C.CDLL(os.path.join(ROOT, "corona-13_b200", "libcorona_host.so"))
    for s in ["accel_print_info", "accel_init", "accel_cleanup", "accel_build", "accel_intersect", "accel_visible",
              "accel_closest", "accel_aabb", "accel_intersect_n", "accel_visible_n",
              "prims_init", "prims_allocate", "prims_load", "prims_allocate_index", "prims_cleanup"]:
        assert hasattr(L, s), s
    if C.CDLL(os.path.join(ROOT, "corona-13_b200", "libcorona_b200.so")).cb200_device_count() < 1:
        L.accel_init.restype = C.c_void_p
        assert L.accel_init(None) is None   # refuses loudly on stderr instead of falling back


def test_host_layer_compiles_against_reference_headers(built, tmp_path):
    """drop-in check: host/accel_b200.c built as the reference's src/accel.d/b200.c would be, with the
    reference's own headers providing ray_t / hit_t / prims_t (only possible where /root/reference exists)"""
    ref = "/root/reference"
    if not os.path.exists(os.path.join(ref, "include", "accel.h")):
        pytest.skip("reference tree not present")
    import subprocess
    out = tmp_path / "b200.o"
    cmd = ["/usr/bin/gcc", "-std=c11", "-D_GNU_SOURCE", "-DCORONA_B200_IN_TREE", "-fno-strict-aliasing", "-c",
           os.path.join(ROOT, "corona-13_b200", "host", "accel_b200.c"), "-o", str(out),
           "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(ROOT, "corona-13_b200", "host"),
           "-I" + os.path.join(ref, "include"), "-I" + ref, "-I" + os.path.join(ref, "ext", "pthread-pool")]
    subprocess.check_call(cmd)
    syms = subprocess.check_output(["nm", str(out)]).decode()
    for s in ["accel_init", "accel_build", "accel_intersect", "accel_visible", "accel_closest", "accel_aabb", "accel_cleanup"]:
        assert f" T {s}" in syms


def test_record_layouts_match_reference_structs(built, tmp_path):
    """sizeof/offsetof of the restated records == the reference's (corona_common.h, prims.h)"""
    ref = "/root/reference"
    if not os.path.exists(os.path.join(ref, "include", "prims.h")):
        pytest.skip("reference tree not present")
    import subprocess
    src = tmp_path / "layout.c"
    src.write_text(r'''
#include <stddef.h>
#include <stdio.h>
#include "corona_common.h"
#include "prims.h"
int main(void) {
  printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\n", sizeof(ray_t), sizeof(hit_t), sizeof(primid_t), sizeof(prims_vtx_t),
    sizeof(prims_vtxidx_t), sizeof(prims_header_t), offsetof(hit_t, dist), offsetof(hit_t, x), offsetof(ray_t, ignore),
    sizeof(prims_shape_t), offsetof(prims_shape_t, vtxidx));
  primid_t p = {5u, 1234567u, 7654321u, 1u, 4u};
  printf("%llu\n", *(unsigned long long *)&p);
  return 0; }''')
    exe = tmp_path / "layout"
    subprocess.check_call(["/usr/bin/gcc", "-std=c11", "-D_GNU_SOURCE", "-fno-strict-aliasing", "-w", str(src), "-o", str(exe),
                           "-I" + os.path.join(ref, "include"), "-I" + ref, "-I" + os.path.join(ref, "ext", "pthread-pool")])
    l1, l2 = subprocess.check_output([str(exe)]).decode().split("\n")[:2]
    from helpers import R
    assert [int(x) for x in l1.split()] == [R.RAY.itemsize, R.HIT.itemsize, 8, R.VTX.itemsize, R.VTXIDX.itemsize, 32,
                                            R.HIT.fields["dist"][1], R.HIT.fields["x"][1], R.RAY.fields["ignore"][1],
                                            1600, 1576]
    assert int(l2) == int(R.primid_make(5, 1234567, np.uint64(7654321), 1, 4))


def test_host_layer_stops_loudly_without_an_accel(built):
    """accel_init returns NULL without a CUDA device and the reference's callers do not check it (src/main.c:344-347): the
    other accel.h entry points must say so and stop, not dereference it; cleanup of nothing is a no-op"""
    import subprocess
    import sys
    code = '''
import ctypes as C
H = C.CDLL(%r)
f = getattr(H, %r); f.argtypes = [C.c_void_p] * %d; f.restype = C.c_void_p
f(*([None] * %d))
print("returned")
'''
    so = os.path.join(ROOT, "corona-13_b200", "libcorona_host.so")
    for name, nargs in (("accel_build", 2), ("accel_aabb", 1)):
        p = subprocess.run([sys.executable, "-c", code % (so, name, nargs, nargs)], capture_output=True, text=True)
        assert p.returncode == -6 and "no cpu fallback" in p.stderr and "returned" not in p.stdout, (name, p.returncode, p.stderr[-200:])
    for name in ("accel_cleanup", "render_cleanup"):
        p = subprocess.run([sys.executable, "-c", code % (so, name, 1, 1)], capture_output=True, text=True)
        assert p.returncode == 0 and "returned" in p.stdout, (name, p.returncode, p.stderr[-200:])
