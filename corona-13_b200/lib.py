"""ctypes binding of the product's C ABI (include/corona_b200.h) for tests and bench.py.

Loading fails loudly when libcorona_b200.so has not been built; every compute call raises
Cb200Error when the library reports an error (no CUDA device, bad arguments, ...).  There is no
CPU path behind any of this.
"""
import ctypes as C
import os

import numpy as np

from .records import RAY, HITREC, QBVH_NODE, CShape

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libcorona_b200.so")

# every symbol include/corona_b200.h declares (checked by tests/test_abi.py)
SYMBOLS = [
    "cb200_version", "cb200_last_error", "cb200_device_count", "cb200_set_device", "cb200_sm_count",
    "cb200_malloc", "cb200_free", "cb200_malloc_host", "cb200_free_host", "cb200_memcpy_h2d",
    "cb200_memcpy_d2h", "cb200_stream_sync",
    "cb200_scene_create", "cb200_scene_destroy", "cb200_scene_num_prims",
    "cb200_accel_build", "cb200_accel_import_qbvh", "cb200_accel_destroy", "cb200_accel_num_nodes",
    "cb200_accel_depth", "cb200_accel_aabb", "cb200_accel_export_qbvh", "cb200_accel_layout",
    "cb200_accel_intersect_n", "cb200_accel_visible_n", "cb200_accel_intersect_dev",
    "cb200_accel_visible_dev", "cb200_accel_intersect_counted", "cb200_launch_count",
]


class Cb200Error(RuntimeError):
    pass


_lib = None


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise Cb200Error(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                         "(there is no CPU fallback)")
    L = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    vp, u64, i32, f32p = C.c_void_p, C.c_uint64, C.c_int, C.POINTER(C.c_float)
    L.cb200_version.restype = C.c_char_p
    L.cb200_last_error.restype = C.c_char_p
    L.cb200_device_count.restype = i32
    L.cb200_set_device.argtypes = [i32]
    L.cb200_sm_count.restype = i32
    L.cb200_malloc.restype = vp
    L.cb200_malloc.argtypes = [C.c_size_t]
    L.cb200_free.argtypes = [vp]
    L.cb200_malloc_host.restype = vp
    L.cb200_malloc_host.argtypes = [C.c_size_t]
    L.cb200_free_host.argtypes = [vp]
    L.cb200_memcpy_h2d.argtypes = [vp, vp, C.c_size_t, vp]
    L.cb200_memcpy_d2h.argtypes = [vp, vp, C.c_size_t, vp]
    L.cb200_stream_sync.argtypes = [vp]
    L.cb200_scene_create.restype = vp
    L.cb200_scene_create.argtypes = [vp, i32]
    L.cb200_scene_destroy.argtypes = [vp]
    L.cb200_scene_num_prims.restype = u64
    L.cb200_scene_num_prims.argtypes = [vp]
    L.cb200_accel_build.restype = vp
    L.cb200_accel_build.argtypes = [vp, vp, vp]
    L.cb200_accel_import_qbvh.restype = vp
    L.cb200_accel_import_qbvh.argtypes = [vp, vp, u64, vp, vp]
    L.cb200_accel_destroy.argtypes = [vp]
    L.cb200_accel_num_nodes.restype = u64
    L.cb200_accel_num_nodes.argtypes = [vp]
    L.cb200_accel_depth.argtypes = [vp]
    L.cb200_accel_aabb.argtypes = [vp, vp]
    L.cb200_accel_export_qbvh.argtypes = [vp, vp, u64, vp]
    L.cb200_accel_layout.argtypes = [vp, vp, vp]
    L.cb200_accel_intersect_n.argtypes = [vp, vp, vp, vp, u64]
    L.cb200_accel_visible_n.argtypes = [vp, vp, vp, vp, u64]
    L.cb200_accel_intersect_dev.argtypes = [vp, vp, vp, vp, u64, vp]
    L.cb200_accel_visible_dev.argtypes = [vp, vp, vp, vp, u64, vp]
    L.cb200_accel_intersect_counted.argtypes = [vp, vp, vp, vp, u64, vp]
    L.cb200_launch_count.restype = u64
    _lib = L
    return L


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _check(rc, what):
    if rc != 0:
        raise Cb200Error(f"{what} failed ({rc}): {load().cb200_last_error().decode()}")


def _nonnull(p, what):
    if not p:
        raise Cb200Error(f"{what} failed: {load().cb200_last_error().decode()}")
    return p


def device_count():
    return load().cb200_device_count()


def set_device(i):
    _check(load().cb200_set_device(i), "cb200_set_device")


def launch_count():
    return int(load().cb200_launch_count())


class Accel:
    """scene upload + acceleration structure on the current device"""

    def __init__(self, scene):
        self.L = load()
        self.host_scene = scene
        self._cshapes = scene.cshapes()
        self.s = _nonnull(self.L.cb200_scene_create(C.cast(self._cshapes, C.c_void_p), len(scene.shapes)),
                          "cb200_scene_create")
        self.a = None
        self.primid = None

    def num_prims(self):
        return int(self.L.cb200_scene_num_prims(self.s))

    def build(self, ghost_aabb=None):
        self._drop_accel()
        n = self.num_prims()
        self.primid = np.zeros(n, np.uint64)
        g = np.ascontiguousarray(ghost_aabb, np.float32) if ghost_aabb is not None else None
        self.a = _nonnull(self.L.cb200_accel_build(self.s, _ptr(g), _ptr(self.primid) if n else None), "cb200_accel_build")
        return self

    def import_qbvh(self, nodes, primid, aabb=None):
        self._drop_accel()
        nodes = np.ascontiguousarray(nodes, dtype=QBVH_NODE)
        primid = np.ascontiguousarray(primid, np.uint64)
        ab = np.ascontiguousarray(aabb, np.float32) if aabb is not None else None
        self.a = _nonnull(self.L.cb200_accel_import_qbvh(self.s, _ptr(nodes), len(nodes), _ptr(primid), _ptr(ab)),
                          "cb200_accel_import_qbvh")
        self.primid = primid.copy()
        return self

    def export_qbvh(self):
        n = int(self.L.cb200_accel_num_nodes(self.a))
        nodes = np.zeros(n, QBVH_NODE)
        primid = np.zeros(self.num_prims(), np.uint64)
        _check(self.L.cb200_accel_export_qbvh(self.a, _ptr(nodes), n, _ptr(primid) if len(primid) else None),
               "cb200_accel_export_qbvh")
        return nodes, primid

    def num_nodes(self):
        return int(self.L.cb200_accel_num_nodes(self.a))

    def depth(self):
        return int(self.L.cb200_accel_depth(self.a))

    def aabb(self):
        out = np.zeros(6, np.float32)
        _check(self.L.cb200_accel_aabb(self.a, _ptr(out)), "cb200_accel_aabb")
        return out

    def layout(self):
        nb, pb = C.c_uint32(0), C.c_uint32(0)
        _check(self.L.cb200_accel_layout(self.a, C.byref(nb), C.byref(pb)), "cb200_accel_layout")
        return nb.value, pb.value

    # host-buffer calls (copies inside)
    def intersect(self, rays, max_dist=None):
        rays = np.ascontiguousarray(rays, dtype=RAY)
        md = np.ascontiguousarray(max_dist, np.float32) if max_dist is not None else None
        out = np.zeros(len(rays), HITREC)
        _check(self.L.cb200_accel_intersect_n(self.a, _ptr(rays), _ptr(md), _ptr(out), len(rays)), "cb200_accel_intersect_n")
        return out

    def visible(self, rays, max_dist):
        rays = np.ascontiguousarray(rays, dtype=RAY)
        md = np.ascontiguousarray(max_dist, np.float32)
        out = np.zeros(len(rays), np.int32)
        _check(self.L.cb200_accel_visible_n(self.a, _ptr(rays), _ptr(md), _ptr(out), len(rays)), "cb200_accel_visible_n")
        return out

    # device-pointer calls (ints from torch .data_ptr(), stream from torch.cuda.current_stream().cuda_stream)
    def intersect_dev(self, d_rays, d_max_dist, d_out, n, stream=0):
        _check(self.L.cb200_accel_intersect_dev(self.a, d_rays, d_max_dist or None, d_out, n, stream or None),
               "cb200_accel_intersect_dev")

    def visible_dev(self, d_rays, d_max_dist, d_out, n, stream=0):
        _check(self.L.cb200_accel_visible_dev(self.a, d_rays, d_max_dist, d_out, n, stream or None),
               "cb200_accel_visible_dev")

    def intersect_counted(self, d_rays, d_max_dist, d_out, n):
        cnt = np.zeros(4, np.uint64)
        _check(self.L.cb200_accel_intersect_counted(self.a, d_rays, d_max_dist or None, d_out, n, _ptr(cnt)),
               "cb200_accel_intersect_counted")
        return cnt

    def _drop_accel(self):
        if self.a:
            self.L.cb200_accel_destroy(self.a)
            self.a = None

    def close(self):
        self._drop_accel()
        if self.s:
            self.L.cb200_scene_destroy(self.s)
            self.s = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
